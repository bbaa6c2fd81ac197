set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/r03r_bench_n8.json 2> gpurun_out/r03r_bench_n8.err
tail -2 gpurun_out/r03r_bench_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r03r_bench_n8.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["ms_per_step"],3), d["parity_ok"], d["e2e"]["ms_per_step"], (d.get("open") or {}).get("ms_per_step"), (d.get("msm_2p24") or {}).get("ms_per_step"), (d.get("mctx") or {}).get("ms_per_step"), (d.get("weak") or {}).get("ms_per_step"))
PY
