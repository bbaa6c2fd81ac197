set -x
for c in 2 4 8 16; do
  python bench.py --cols $c --no-cpu --steps 5 --warmup 3 > gpurun_out/r02a_bench_cols$c.json 2> gpurun_out/r02a_bench_cols$c.err
done
python bench.py --workload msm --log-n 20 --msm-cols 1 --steps 5 --warmup 3 > gpurun_out/r02a_msm_2p20x1.json 2>&1
python bench.py --workload msm --log-n 21 --msm-cols 1 --steps 5 --warmup 3 > gpurun_out/r02a_msm_2p21x1.json 2>&1
python bench.py --workload msm --log-n 24 --msm-cols 1 --steps 3 --warmup 2 > gpurun_out/r02a_msm_2p24x1.json 2>&1
python tools/copy2d_probe.py > gpurun_out/r02a_copy2d.log 2>&1
nvidia-smi topo -m > gpurun_out/r02a_topo.txt 2>&1
lscpu | head -30 > gpurun_out/r02a_lscpu.txt
