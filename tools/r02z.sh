set -x
python -m pytest tests/test_gpu_msm_rounds.py -m gpu -x -q -k "fused or dlog or skewed" > gpurun_out/r02z_pytest.log 2>&1
tail -3 gpurun_out/r02z_pytest.log
B="python bench.py --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx"
$B > gpurun_out/r02z_bench_c16.json 2> gpurun_out/r02z_bench_c16.err
for sp in 0 1; do for c in 2 4; do
EON_MSM_SPLIT=$sp $B --cols $c > gpurun_out/r02z_bench_split${sp}_c$c.json 2> gpurun_out/r02z_bench_split${sp}_c$c.err
done; done
python - <<'PY'
import json
for f in ("c16","split0_c2","split1_c2","split0_c4","split1_c4"):
    try:
        d=json.loads(open(f"gpurun_out/r02z_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["two_calls"]["ms_per_step"],3), d["parity_ok"])
        print({k:round(v,2) for k,v in d["phase_ms_per_step"].items() if v})
    except Exception as e:
        print(f, "failed", e)
PY
