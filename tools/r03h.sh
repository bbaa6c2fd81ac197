set -x
B="python bench.py --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx"
$B > gpurun_out/r03h_bench_default.json 2> gpurun_out/r03h_bench_default.err
cp plonky3_eon_b200/libeon_kzg.so /tmp/libeon_default.so
cp plonky3_eon_b200/libeon_kzg_sqr.so plonky3_eon_b200/libeon_kzg.so
$B > gpurun_out/r03h_bench_sqr.json 2> gpurun_out/r03h_bench_sqr.err
cp /tmp/libeon_default.so plonky3_eon_b200/libeon_kzg.so
python - <<'PY'
import json
for f in ("default","sqr"):
    try:
        d=json.loads(open(f"gpurun_out/r03h_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["two_calls"]["ms_per_step"],3), d["parity_ok"])
        print({k:round(v,2) for k,v in d["phase_ms_per_step"].items() if v})
    except Exception as ex:
        print(f, "failed", ex)
PY
