set -x
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx"
$CMD > gpurun_out/r03s_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r03s_ncu_launches.csv $CMD > gpurun_out/r03s_ncu1.log 2>&1
