#!/bin/bash
# the velocity BiCGStab with the shipped kernel rule (one resident wave of tiled CTAs) and with the row-per-thread kernels, same box
mkdir -p gpurun_out
timeout 40 python scripts/velocity_bench.py --size 128 128 128 --no-cpu --no-csr --reps 3 --tiles -1 0 > gpurun_out/r02E_vel.log 2>&1
grep -h iterations_per_s gpurun_out/r02E_vel.log | cut -c1-330
