#!/bin/bash
# ncu --set full of one launch of the tiled BiCGStab SpMV kernels (128^3 velocity system) + two more chunk sizes
set -u
mkdir -p gpurun_out
timeout 150 ncu --set full --import-source on --clock-control none -k regex:k_sep_tile_bcgs_spmv -s 8 -c 2 -o gpurun_out/r02B_sep_tile -f \
  python scripts/velocity_bench.py --size 128 128 128 --no-cpu --no-csr --reps 1 --iters 8 --tiles 2 > gpurun_out/r02B_ncu.log 2>&1
echo "ncu exit $?" | tee -a gpurun_out/r02B_summary.log
timeout 60 python scripts/velocity_bench.py --size 128 128 128 --no-cpu --no-csr --reps 2 --tiles 2 --zchunks 4 6 > gpurun_out/r02B_vel.log 2>&1
grep -h iterations_per_s gpurun_out/r02B_vel.log | cut -c1-400 | tee -a gpurun_out/r02B_summary.log
