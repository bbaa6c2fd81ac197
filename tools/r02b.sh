set -x
python -m pytest tests/test_gpu_mctx.py tests/test_gpu_kzg.py -m gpu -x -q > gpurun_out/r02b_pytest_new.log 2>&1
tail -5 gpurun_out/r02b_pytest_new.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
tail -3 gpurun_out/r02b_bench.err
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02b_pytest_gpu.log
