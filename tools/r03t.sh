set -x
python -m pytest tests/test_gpu_kzg.py tests/test_gpu_msm_rounds.py -m gpu -x -q > gpurun_out/r03t_pytest.log 2>&1
tail -2 gpurun_out/r03t_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03t_smoke.log 2>&1
tail -1 gpurun_out/r03t_smoke.log
python bench.py --no-cpu --no-open --msm-log-n 0 --no-mctx > gpurun_out/r03t_bench.json 2> gpurun_out/r03t_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r03t_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],3), round(d["two_calls"]["ms_per_step"],3), d["parity_ok"], d["e2e"]["ms_per_step"])
PY
