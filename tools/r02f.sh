set -x
python -m pytest tests/test_gpu_msm_rounds.py tests/test_gpu_kzg.py -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1
tail -4 gpurun_out/r02f_pytest.log
B="python bench.py --no-cpu --msm-log-n 0 --no-mctx --no-open"
$B > gpurun_out/r02f_n1.json 2> gpurun_out/r02f_n1.err
$B --cols 2 > gpurun_out/r02f_c2_default.json 2> gpurun_out/r02f_c2_default.err
EON_TREE_B=16 $B --cols 2 > gpurun_out/r02f_c2_treeb16.json 2> gpurun_out/r02f_c2_treeb16.err
EON_TREE_B=8 $B --cols 2 > gpurun_out/r02f_c2_treeb8.json 2> gpurun_out/r02f_c2_treeb8.err
EON_MSM_SPLIT=0 EON_TREE_SLICED_B=8 $B --cols 2 > gpurun_out/r02f_c2_nosplit_slb8.json 2> gpurun_out/r02f_c2_nosplit_slb8.err
EON_MSM_SPLIT=0 EON_TREE_SLICED_B=16 $B --cols 2 > gpurun_out/r02f_c2_nosplit_slb16.json 2> gpurun_out/r02f_c2_nosplit_slb16.err
EON_MSM_SPLIT=0 $B --cols 2 > gpurun_out/r02f_c2_nosplit.json 2> gpurun_out/r02f_c2_nosplit.err
EON_NTT_MINB=3 $B --cols 2 > gpurun_out/r02f_c2_nttminb3.json 2> gpurun_out/r02f_c2_nttminb3.err
$B --cols 4 > gpurun_out/r02f_c4_default.json 2> gpurun_out/r02f_c4_default.err
$B --cols 8 > gpurun_out/r02f_c8_default.json 2> gpurun_out/r02f_c8_default.err
