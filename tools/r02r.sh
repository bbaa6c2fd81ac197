set -x
python -m pytest tests/test_gpu_dft.py tests/test_gpu_ntt_forms.py tests/test_gpu_msm_rounds.py tests/test_gpu_kzg.py -m gpu -x -q > gpurun_out/r02r_pytest.log 2>&1
tail -5 gpurun_out/r02r_pytest.log
B="python bench.py --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx"
for db in 0 1 2; do
EON_NTT_DB=$db $B > gpurun_out/r02r_bench_db$db.json 2> gpurun_out/r02r_bench_db$db.err
done
python - <<'PY'
import json
for db in (0,1,2):
    try:
        d=json.loads(open(f"gpurun_out/r02r_bench_db{db}.json").read().strip().splitlines()[-1])
        p=d["phase_ms_per_step"]
        print(db, d["ms_per_step"], d["parity_ok"], {k:round(v,2) for k,v in p.items()})
    except Exception as e:
        print(db, "failed", e)
PY
