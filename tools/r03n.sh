set -x
python -m pytest tests/test_gpu_kzg.py tests/test_gpu_msm_rounds.py tests/test_gpu_mctx.py -m gpu -x -q > gpurun_out/r03n_pytest.log 2>&1
tail -3 gpurun_out/r03n_pytest.log
B="python bench.py --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx"
$B --cols 2 > gpurun_out/r03n_bench_c2.json 2> gpurun_out/r03n_bench_c2.err
$B > gpurun_out/r03n_bench_c16.json 2> gpurun_out/r03n_bench_c16.err
python - <<'PY'
import json
for f in ("c2","c16"):
    try:
        d=json.loads(open(f"gpurun_out/r03n_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["two_calls"]["ms_per_step"],3), d["parity_ok"], {k:round(v,2) for k,v in d["phase_ms_per_step"].items() if v})
    except Exception as ex:
        print(f, "failed", ex)
PY
