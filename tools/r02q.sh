set -x
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx"
python bench.py --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx > gpurun_out/r02q_bench_short.json 2> gpurun_out/r02q_bench_short.err
$CMD > gpurun_out/r02q_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_bin_hist|k_sort_coarse|k_sort_fine|k_pair_hist|k_pair_scatter|k_tree_fwd|k_msm_accumulate|k_bucket_rowcol|k_sort_count" -s 11 -c 11 -f -o /tmp/r02q_mem $CMD > gpurun_out/r02q_ncu.log 2>&1
ls -la /tmp/*.ncu-rep
ncu -i /tmp/r02q_mem.ncu-rep --page raw --csv > gpurun_out/r02q_mem_raw.csv
ncu -i /tmp/r02q_mem.ncu-rep --page source --csv -k regex:k_sort_coarse > gpurun_out/r02q_src_coarse.csv
ncu -i /tmp/r02q_mem.ncu-rep --page source --csv -k regex:k_sort_fine > gpurun_out/r02q_src_fine.csv
ncu -i /tmp/r02q_mem.ncu-rep --page source --csv -k regex:k_pair_scatter > gpurun_out/r02q_src_pair_scatter.csv
ncu -i /tmp/r02q_mem.ncu-rep --page source --csv -k regex:k_tree_fwd_sliced > gpurun_out/r02q_src_fwd_sliced.csv
ncu -i /tmp/r02q_mem.ncu-rep --page source --csv -k regex:k_bin_hist > gpurun_out/r02q_src_bin_hist.csv
sz=$(stat -c %s /tmp/r02q_mem.ncu-rep); if [ $sz -lt 30000000 ]; then cp /tmp/r02q_mem.ncu-rep gpurun_out/; fi
du -sh gpurun_out
