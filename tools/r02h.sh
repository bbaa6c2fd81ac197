set -x
python -m pytest tests/test_gpu_dft.py tests/test_gpu_large.py tests/test_gpu_mctx.py tests/test_gpu_msm_rounds.py -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1
tail -6 gpurun_out/r02h_pytest.log
B="python bench.py --no-cpu --msm-log-n 0 --no-e2e --no-open"
for st in 0 1; do for c in 2 4; do
EON_MSM_STAGGER=$st $B --cols $c > gpurun_out/r02h_stagger${st}_c$c.json 2> gpurun_out/r02h_stagger${st}_c$c.err
done; done
python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-open --msm-log-n 0"
$CMD > gpurun_out/r02h_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02h_ncu_launches.csv $CMD > gpurun_out/r02h_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 27 -c 6 -f -o gpurun_out/r02h_ntt $CMD > gpurun_out/r02h_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tree_bwd -s 12 -c 3 -f -o gpurun_out/r02h_bwd $CMD > gpurun_out/r02h_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
