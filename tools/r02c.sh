set -x
nvidia-smi topo -m > gpurun_out/r02c_topo.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02c_pytest_gpu.log
python bench.py --cols 2 --no-cpu --msm-log-n 21 > gpurun_out/r02c_bench_cols2.json 2> gpurun_out/r02c_bench_cols2.err
python bench.py --no-cpu --msm-log-n 0 --no-open > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err
tail -3 gpurun_out/r02c_bench_n2.err
