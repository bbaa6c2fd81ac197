set -x
python -m pytest tests/test_gpu_dft.py tests/test_gpu_mctx.py -m gpu -x -q > gpurun_out/r03b_pytest.log 2>&1
tail -3 gpurun_out/r03b_pytest.log
B="python bench.py --no-cpu --no-open --msm-log-n 0 --no-mctx"
for v in 0 1; do
EON_NTT_WIDEN=$v $B > gpurun_out/r03b_bench_w$v.json 2> gpurun_out/r03b_bench_w$v.err
EON_NTT_WIDEN=$v $B --cols 2 --no-e2e > gpurun_out/r03b_bench_w${v}_c2.json 2> gpurun_out/r03b_bench_w${v}_c2.err
EON_NTT_WIDEN=$v $B --cols 4 --no-e2e > gpurun_out/r03b_bench_w${v}_c4.json 2> gpurun_out/r03b_bench_w${v}_c4.err
done
python - <<'PY'
import json
for f in ("w0","w1","w0_c2","w1_c2","w0_c4","w1_c4"):
    try:
        d=json.loads(open(f"gpurun_out/r03b_bench_{f}.json").read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f, round(d["ms_per_step"],3), round(d["two_calls"]["ms_per_step"],3), d["parity_ok"], e.get("ms_per_step"), e.get("two_calls_ms_per_step"), round(d["phase_ms_per_step"]["ntt_passes"],3), (e.get("phase_ms_per_step") or {}).get("ntt_passes"))
    except Exception as ex:
        print(f, "failed", ex)
PY
