set -x
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx"
$CMD > gpurun_out/r02x_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_sort_coarse|k_sort_place2|k_sort_count2" -s 6 -c 3 -f -o /tmp/r02x_sort $CMD > gpurun_out/r02x_ncu.log 2>&1
ls -la /tmp/*.ncu-rep
ncu -i /tmp/r02x_sort.ncu-rep --page raw --csv > gpurun_out/r02x_sort_raw.csv
ncu -i /tmp/r02x_sort.ncu-rep --page source --csv -k regex:k_sort_place2 > gpurun_out/r02x_src_place2.csv
ncu -i /tmp/r02x_sort.ncu-rep --page source --csv -k regex:k_sort_coarse > gpurun_out/r02x_src_coarse.csv
du -sh gpurun_out
