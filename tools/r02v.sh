set -x
B="python bench.py --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx"
for f in 7 6; do
EON_SORT_FB=$f $B > gpurun_out/r02v_bench_fb$f.json 2> gpurun_out/r02v_bench_fb$f.err
done
EON_SORT_FB=6 EON_SORT_FUSED=0 $B > gpurun_out/r02v_bench_fb6_unfused.json 2> gpurun_out/r02v_bench_fb6_unfused.err
python - <<'PY'
import json
for f in ("fb7","fb6","fb6_unfused"):
    try:
        d=json.loads(open(f"gpurun_out/r02v_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["parity_ok"])
        print({k:round(v,2) for k,v in d["phase_ms_per_step"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
