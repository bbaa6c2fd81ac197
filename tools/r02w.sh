set -x
B="python bench.py --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx --check-e2e"
for v in 0 1; do
EON_LDE_AFTER_R0=$v $B > gpurun_out/r02w_bench_r0$v.json 2> gpurun_out/r02w_bench_r0$v.err
EON_LDE_AFTER_R0=$v $B --cols 2 > gpurun_out/r02w_bench_r0${v}_c2.json 2> gpurun_out/r02w_bench_r0${v}_c2.err
EON_LDE_AFTER_R0=$v $B --cols 4 > gpurun_out/r02w_bench_r0${v}_c4.json 2> gpurun_out/r02w_bench_r0${v}_c4.err
done
python - <<'PY'
import json
for f in ("r00","r01","r00_c2","r01_c2","r00_c4","r01_c4"):
    try:
        d=json.loads(open(f"gpurun_out/r02w_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["two_calls"]["ms_per_step"],3), d["parity_ok"])
    except Exception as e:
        print(f, "failed", e)
PY
