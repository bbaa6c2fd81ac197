set -x
python -m pytest tests/test_gpu_msm_rounds.py tests/test_gpu_kzg.py tests/test_gpu_batched_pcs.py tests/test_gpu_mctx.py tests/test_gpu_large.py -m gpu -x -q > gpurun_out/r02u_pytest.log 2>&1
tail -15 gpurun_out/r02u_pytest.log
B="python bench.py --no-cpu --no-open --msm-log-n 0 --no-mctx"
for f in 0 1; do
EON_SORT_FUSED=$f $B > gpurun_out/r02u_bench_fused$f.json 2> gpurun_out/r02u_bench_fused$f.err
done
python - <<'PY'
import json
for f in (0,1):
    try:
        d=json.loads(open(f"gpurun_out/r02u_bench_fused{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["parity_ok"], d["e2e"]["ms_per_step"], d["e2e"].get("two_calls_ms_per_step"))
        print({k:round(v,2) for k,v in d["phase_ms_per_step"].items()})
        print({k:round(v,2) for k,v in d["e2e"]["phase_ms_per_step"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
