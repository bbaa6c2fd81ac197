# 8 x B200: strong scaling of the headline config, BASELINE configs[4], index-range MSMs, the multi-device context
set -x
nvidia-smi topo -m > gpurun_out/r02i_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29501 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02i_bench_n8.json 2> gpurun_out/r02i_bench_n8.err
tail -2 gpurun_out/r02i_bench_n8.err
$TR --nproc-per-node 4 --master-port 29502 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02i_bench_n4.json 2> gpurun_out/r02i_bench_n4.err
$TR --nproc-per-node 8 --master-port 29503 bench.py --gpus 8 --log-rows 24 --cols 64 --added-bits 2 --no-e2e --no-weak --no-open --msm-log-n 0 --steps 2 --warmup 1 > gpurun_out/r02i_cfg5_n8.json 2> gpurun_out/r02i_cfg5_n8.err
tail -2 gpurun_out/r02i_cfg5_n8.err
$TR --nproc-per-node 8 --master-port 29504 bench.py --gpus 8 --workload msm --log-n 24 --msm-cols 8 --steps 3 --warmup 2 > gpurun_out/r02i_msm24x8_n8.json 2> gpurun_out/r02i_msm24x8_n8.err
$TR --nproc-per-node 8 --master-port 29505 bench.py --gpus 8 --workload msm --log-n 26 --msm-cols 1 --steps 3 --warmup 2 > gpurun_out/r02i_msm26_n8.json 2> gpurun_out/r02i_msm26_n8.err
python -m pytest tests/test_gpu_mctx.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02i_pytest_multi.log 2>&1
tail -3 gpurun_out/r02i_pytest_multi.log
python bench.py --workload prove-pcs --mctx-devices 8 --steps 3 --warmup 2 > gpurun_out/r02i_prove_pcs_mctx8.json 2> gpurun_out/r02i_prove_pcs_mctx8.err
python bench.py --workload prove-pcs --steps 3 --warmup 2 > gpurun_out/r02i_prove_pcs_1gpu.json 2> gpurun_out/r02i_prove_pcs_1gpu.err
