set -x
CMD="python bench.py --cols 2 --steps 1 --warmup 1 --no-cpu --no-e2e --no-open --msm-log-n 0 --no-mctx"
$CMD > gpurun_out/r03f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r03f_ncu_launches_c2.csv $CMD > gpurun_out/r03f_ncu1.log 2>&1
