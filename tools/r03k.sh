set -x
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx"
$CMD > gpurun_out/r03k_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r03k_ncu_launches.csv $CMD > gpurun_out/r03k_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_tree_bwd|k_ntt_pass" -s 18 -c 9 -f -o /tmp/r03k_top $CMD > gpurun_out/r03k_ncu2.log 2>&1
python tools/ncu_summary.py /tmp/r03k_top.ncu-rep > gpurun_out/r03k_ncu_full_top.txt 2> gpurun_out/r03k_sum1.err
ncu --set full --clock-control none --import-source on -k "regex:k_sort_coarse|k_sort_place2|k_sort_count2|k_bin_hist|k_tree_fwd_sliced|k_msm_accumulate<\(bool\)1>|k_bucket_rowcol|k_tree_top" -s 16 -c 8 -f -o /tmp/r03k_rest $CMD > gpurun_out/r03k_ncu3.log 2>&1
python tools/ncu_summary.py /tmp/r03k_rest.ncu-rep > gpurun_out/r03k_ncu_full_rest.txt 2> gpurun_out/r03k_sum2.err
ls -la /tmp/*.ncu-rep
cuobjdump -sass plonky3_eon_b200/csrc/msm_tree.o 2>/dev/null | awk '/Function :/{name=$3} /^ +\/\*[0-9a-f]+\*\/ /{op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); cnt[name" "op]++} END{for(k in cnt) print cnt[k], k}' | grep "k_tree_bwd_slicedILi32" | sort -rn | head -40 > gpurun_out/r03k_sass_k_tree_bwd_sliced.txt
du -sh gpurun_out
