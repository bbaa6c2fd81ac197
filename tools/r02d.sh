set -x
for prio in 0 1; do for c in 2 16; do
EON_MSM_PRIO=$prio python bench.py --cols $c --no-cpu --msm-log-n 0 --no-open --no-mctx > gpurun_out/r02d_prio${prio}_cols$c.json 2> gpurun_out/r02d_prio${prio}_cols$c.err
done; done
for c in 2 4; do
python bench.py --cols $c --no-cpu --msm-log-n 0 --no-open --no-mctx --slice-schedule 1 > gpurun_out/r02d_slice1_cols$c.json 2> gpurun_out/r02d_slice1_cols$c.err
done
for wb in 16 17 18 19 20; do
python bench.py --workload msm --log-n 21 --window-bits $wb > gpurun_out/r02d_msm21_c$wb.json 2> gpurun_out/r02d_msm21_c$wb.err
done
for wb in 18 19; do
python bench.py --workload msm --log-n 24 --window-bits $wb --steps 3 > gpurun_out/r02d_msm24_c$wb.json 2> gpurun_out/r02d_msm24_c$wb.err
done
