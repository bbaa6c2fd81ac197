set -x
python -m pytest tests/test_gpu_kzg.py tests/test_gpu_msm_rounds.py tests/test_gpu_batched_pcs.py tests/test_gpu_mctx.py tests/test_gpu_large.py -m gpu -x -q > gpurun_out/r03i_pytest.log 2>&1
tail -3 gpurun_out/r03i_pytest.log
B="python bench.py --no-cpu --no-e2e --msm-log-n 24 --no-mctx"
$B > gpurun_out/r03i_bench_c16.json 2> gpurun_out/r03i_bench_c16.err
python - <<'PY'
import json
for f in ("c16",):
    try:
        d=json.loads(open(f"gpurun_out/r03i_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["two_calls"]["ms_per_step"],3), d["parity_ok"], d["open"]["ms_per_step"], d["msm_2p24"]["ms_per_step"])
        print({k:round(v,2) for k,v in d["phase_ms_per_step"].items() if v})
    except Exception as ex:
        print(f, "failed", ex)
PY
