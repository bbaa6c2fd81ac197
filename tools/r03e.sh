set -x
python -m pytest tests/test_gpu_multi.py tests/test_gpu_mctx.py -m gpu -x -q > gpurun_out/r03e_pytest_2gpu.log 2>&1
tail -3 gpurun_out/r03e_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r03e_bench_n2.json 2> gpurun_out/r03e_bench_n2.err
tail -3 gpurun_out/r03e_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r03e_bench_ref_n2.json 2> gpurun_out/r03e_bench_ref_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r03e_bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["ms_per_step"], d["parity_ok"], d["e2e"]["ms_per_step"], d["open"]["ms_per_step"], d["msm_2p24"]["ms_per_step"], d["mctx"]["ms_per_step"], d["clocks"])
print(d["parity"])
PY
