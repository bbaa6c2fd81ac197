set -x
python -m pytest tests/test_gpu_kzg.py tests/test_gpu_mctx.py tests/test_gpu_batched_pcs.py -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1
tail -3 gpurun_out/r02k_pytest.log
B="python bench.py --no-cpu --msm-log-n 0 --no-open"
$B > gpurun_out/r02k_new_c16.json 2> gpurun_out/r02k_new_c16.err
EON_PIPE_MSM_PER_GROUP=1 $B > gpurun_out/r02k_old_c16.json 2> gpurun_out/r02k_old_c16.err
$B --cols 8 > gpurun_out/r02k_new_c8.json 2> gpurun_out/r02k_new_c8.err
EON_PIPE_MSM_PER_GROUP=1 $B --cols 8 > gpurun_out/r02k_old_c8.json 2> gpurun_out/r02k_old_c8.err
python bench.py --workload prove-pcs --steps 3 --warmup 2 > gpurun_out/r02k_prove_pcs.json 2> gpurun_out/r02k_prove_pcs.err
