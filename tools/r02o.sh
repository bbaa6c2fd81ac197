set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02o_bench_n8.json 2> gpurun_out/r02o_bench_n8.err
tail -2 gpurun_out/r02o_bench_n8.err
$TR --nproc-per-node 4 --master-port 29532 bench.py --gpus 4 --steps 5 --warmup 3 --msm-log-n 0 --no-open --no-weak > gpurun_out/r02o_bench_n4.json 2> gpurun_out/r02o_bench_n4.err
