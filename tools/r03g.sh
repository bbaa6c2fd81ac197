set -x
B="python bench.py --no-cpu --no-open --msm-log-n 0 --no-mctx"
for v in 0 1; do
EON_PIPE_FIRST2=$v $B > gpurun_out/r03g_bench_f$v.json 2> gpurun_out/r03g_bench_f$v.err
done
python - <<'PY'
import json
for f in ("f0","f1"):
    try:
        d=json.loads(open(f"gpurun_out/r03g_bench_{f}.json").read().strip().splitlines()[-1])
        e=d["e2e"]
        print(f, round(d["ms_per_step"],3), d["parity_ok"], e["ms_per_step"], e["two_calls_ms_per_step"])
    except Exception as ex:
        print(f, "failed", ex)
PY
