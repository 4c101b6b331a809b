#!/bin/bash
# ncu --set full of the tiled BiCGStab SpMV kernels with the cp.async queue (128^3 velocity system), then the solve with the default rule
set -u
mkdir -p gpurun_out
timeout 120 ncu --set full --import-source on --clock-control none -k regex:k_sep_tile_bcgs_spmv -s 8 -c 2 -o gpurun_out/r02D_sep_tile -f \
  python scripts/velocity_bench.py --size 128 128 128 --no-cpu --no-csr --reps 1 --iters 8 --tiles -1 > gpurun_out/r02D_ncu.log 2>&1
echo "ncu exit $?" | tee -a gpurun_out/r02D_summary.log
timeout 60 python scripts/velocity_bench.py --size 128 128 128 --no-cpu --no-csr --reps 3 --tiles -1 > gpurun_out/r02D_vel.log 2>&1
grep -h iterations_per_s gpurun_out/r02D_vel.log | cut -c1-330 | tee -a gpurun_out/r02D_summary.log
