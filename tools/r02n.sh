set -x
python -m pytest tests/test_gpu_msm_rounds.py tests/test_gpu_mctx.py -m gpu -x -q > gpurun_out/r02n_pytest.log 2>&1
tail -3 gpurun_out/r02n_pytest.log
B="python bench.py --no-cpu --msm-log-n 0 --no-e2e --no-open"
for t in 0 1; do for c in 2 4; do
EON_MSM_TINY_PRIO=$t $B --cols $c > gpurun_out/r02n_tiny${t}_c$c.json 2> gpurun_out/r02n_tiny${t}_c$c.err
done; done
$B > gpurun_out/r02n_c16.json 2> gpurun_out/r02n_c16.err
