set -x
python -m pytest tests/test_gpu_msm_rounds.py tests/test_gpu_kzg.py tests/test_gpu_mctx.py -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1
tail -4 gpurun_out/r02e_pytest.log
for sp in 0 1; do for c in 2 4; do
EON_MSM_SPLIT=$sp python bench.py --cols $c --no-cpu --msm-log-n 0 --no-mctx > gpurun_out/r02e_split${sp}_cols$c.json 2> gpurun_out/r02e_split${sp}_cols$c.err
done; done
python bench.py --no-cpu --msm-log-n 21 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err
