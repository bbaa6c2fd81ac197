set -x
python -m pytest tests/test_gpu_msm_rounds.py -m gpu -x -q > gpurun_out/r02y_pytest.log 2>&1
tail -3 gpurun_out/r02y_pytest.log
B="python bench.py --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx"
for t in 1024 768 512; do
EON_SORT_TILE=$t $B > gpurun_out/r02y_bench_tile$t.json 2> gpurun_out/r02y_bench_tile$t.err
done
python - <<'PY'
import json
for f in ("tile1024","tile768","tile512"):
    try:
        d=json.loads(open(f"gpurun_out/r02y_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["two_calls"]["ms_per_step"],3), d["parity_ok"])
        print({k:round(v,2) for k,v in d["phase_ms_per_step"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
