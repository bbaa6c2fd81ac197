set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r03o_pytest_gpu.log 2>&1
tail -5 gpurun_out/r03o_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03o_smoke.log 2>&1
tail -2 gpurun_out/r03o_smoke.log
python bench.py > gpurun_out/r03o_bench.json 2> gpurun_out/r03o_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r03o_bench_reference.json 2> gpurun_out/r03o_bench_reference.err
