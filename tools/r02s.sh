set -x
python -m pytest tests/test_gpu_kzg.py tests/test_gpu_batched_pcs.py tests/test_gpu_mctx.py tests/test_gpu_msm_rounds.py -m gpu -x -q > gpurun_out/r02s_pytest.log 2>&1
tail -5 gpurun_out/r02s_pytest.log
B="python bench.py --no-cpu --no-open --msm-log-n 0 --no-mctx"
for es in 0 1; do
EON_PIPE_EARLY_SORT=$es $B > gpurun_out/r02s_bench_es$es.json 2> gpurun_out/r02s_bench_es$es.err
done
python - <<'PY'
import json
for es in (0,1):
    try:
        d=json.loads(open(f"gpurun_out/r02s_bench_es{es}.json").read().strip().splitlines()[-1])
        print(es, d["ms_per_step"], d["parity_ok"], d["e2e"]["ms_per_step"], d["e2e"].get("two_calls_ms_per_step"))
    except Exception as e:
        print(es, "failed", e)
PY
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx"
$CMD > gpurun_out/r02s_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02s_ncu_launches.csv $CMD > gpurun_out/r02s_ncu1.log 2>&1
