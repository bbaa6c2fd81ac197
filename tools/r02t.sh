set -x
B="python bench.py --no-cpu --no-open --msm-log-n 0 --no-mctx"
for v in 1 0; do
EON_PIPE_D2H_EARLY=$v $B > gpurun_out/r02t_bench_d2hearly$v.json 2> gpurun_out/r02t_bench_d2hearly$v.err
done
python - <<'PY'
import json
for v in (1,0):
    try:
        d=json.loads(open(f"gpurun_out/r02t_bench_d2hearly{v}.json").read().strip().splitlines()[-1])
        print(v, d["ms_per_step"], d["parity_ok"], d["e2e"]["ms_per_step"], d["e2e"].get("two_calls_ms_per_step"))
    except Exception as e:
        print(v, "failed", e)
PY
