set -x
python -m pytest tests/test_gpu_dft.py tests/test_gpu_ntt_forms.py tests/test_gpu_msm_rounds.py -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1
tail -4 gpurun_out/r02g_pytest.log
B="python bench.py --no-cpu --msm-log-n 0 --no-e2e --no-open"
for c in 2 4 8 16; do for b in 8 16 32; do
EON_TREE_B=$b EON_TREE_SLICED_B=$b $B --cols $c > gpurun_out/r02g_c${c}_b$b.json 2> gpurun_out/r02g_c${c}_b$b.err
done; done
for tws in 0 1; do for mb in 2 3; do
EON_NTT_TWS=$tws EON_NTT_MINB=$mb $B > gpurun_out/r02g_ntt_tws${tws}_minb$mb.json 2> gpurun_out/r02g_ntt_tws${tws}_minb$mb.err
done; done
