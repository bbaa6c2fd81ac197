set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --rowblock-max-cols 8 --msm-log-n 0 --no-open --no-weak --no-mctx > gpurun_out/r02m_bench_n2_rowblock.json 2> gpurun_out/r02m_bench_n2_rowblock.err
tail -5 gpurun_out/r02m_bench_n2_rowblock.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --cols 4 --rowblock-max-cols 8 --msm-log-n 0 --no-open --no-weak --no-mctx > gpurun_out/r02m_bench_n2_c4_rowblock.json 2> gpurun_out/r02m_bench_n2_c4_rowblock.err
python -m pytest tests/test_cxx_mirror.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1
tail -3 gpurun_out/r02m_pytest.log
