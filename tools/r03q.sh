set -x
python bench.py --log-rows 24 --cols 8 --added-bits 2 --no-e2e --no-open --msm-log-n 0 --no-cpu --no-mctx --steps 2 --warmup 1 > gpurun_out/r03q_cfg5shard.json 2> gpurun_out/r03q_cfg5shard.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r03q_cfg5shard.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["parity_ok"], {k:round(v,1) for k,v in d["phase_ms_per_step"].items() if v})
PY
