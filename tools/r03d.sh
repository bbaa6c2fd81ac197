set -x
B="python bench.py --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx"
$B --cols 2 > gpurun_out/r03d_bench_c2.json 2> gpurun_out/r03d_bench_c2.err
$B --cols 4 > gpurun_out/r03d_bench_c4.json 2> gpurun_out/r03d_bench_c4.err
$B > gpurun_out/r03d_bench_c16.json 2> gpurun_out/r03d_bench_c16.err
python - <<'PY'
import json
for f in ("c2","c4","c16"):
    try:
        d=json.loads(open(f"gpurun_out/r03d_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["two_calls"]["ms_per_step"],3), d["parity_ok"])
        print({k:round(v,2) for k,v in d["phase_ms_per_step"].items() if v})
    except Exception as ex:
        print(f, "failed", ex)
PY
