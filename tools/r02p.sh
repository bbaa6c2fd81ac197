set -x
python -m pytest tests/test_gpu_mctx.py tests/test_gpu_batched_pcs.py -m gpu -x -q > gpurun_out/r02p_pytest.log 2>&1
tail -3 gpurun_out/r02p_pytest.log
B="python bench.py --no-cpu --msm-log-n 0 --no-e2e"
EON_MSM_BUDGET_GB=24 $B > gpurun_out/r02p_open_b24.json 2> gpurun_out/r02p_open_b24.err
$B > gpurun_out/r02p_open_b64.json 2> gpurun_out/r02p_open_b64.err
C5="python bench.py --log-rows 24 --cols 8 --added-bits 2 --no-e2e --no-open --msm-log-n 0 --no-cpu --steps 2 --warmup 1"
EON_MSM_BUDGET_GB=24 $C5 > gpurun_out/r02p_cfg5shard_b24.json 2> gpurun_out/r02p_cfg5shard_b24.err
$C5 > gpurun_out/r02p_cfg5shard_b64.json 2> gpurun_out/r02p_cfg5shard_b64.err
