set -x
python -m pytest tests/test_gpu_kzg.py tests/test_gpu_msm_rounds.py -m gpu -x -q > gpurun_out/r03m_pytest.log 2>&1
tail -3 gpurun_out/r03m_pytest.log
B="python bench.py --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx"
for v in 0 1; do
EON_BUCKET_W2=$v $B --cols 2 > gpurun_out/r03m_bench_w2${v}_c2.json 2> gpurun_out/r03m_bench_w2${v}_c2.err
EON_BUCKET_W2=$v $B > gpurun_out/r03m_bench_w2${v}_c16.json 2> gpurun_out/r03m_bench_w2${v}_c16.err
done
python - <<'PY'
import json
for f in ("w20_c2","w21_c2","w20_c16","w21_c16"):
    try:
        d=json.loads(open(f"gpurun_out/r03m_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["two_calls"]["ms_per_step"],3), d["parity_ok"], round(d["phase_ms_per_step"]["msm_reduce"],3))
    except Exception as ex:
        print(f, "failed", ex)
PY
