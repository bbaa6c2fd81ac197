set -x
: > gpurun_out/r03p_msm_sweep_1gpu.jsonl
for L in 16 18 20 22 24 26; do
python bench.py --workload msm --log-n $L --steps 3 --warmup 2 --no-cpu >> gpurun_out/r03p_msm_sweep_1gpu.jsonl 2>> gpurun_out/r03p_msm_sweep.err
done
python - <<'PY'
import json
for l in open("gpurun_out/r03p_msm_sweep_1gpu.jsonl"):
    l=l.strip()
    if not l.startswith("{"): continue
    d=json.loads(l); print(d["config"].get("log_n"), round(d["ms_per_step"],3), "%.3e"%d["value"], d.get("parity_ok"))
PY
