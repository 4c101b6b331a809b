#!/bin/bash
# ncu --set full of the final default kernels (k_spmv4 64 x 4 tiles / 64 planes per CTA, k_update2) at 256^3; CSV pages only
set -u
mkdir -p gpurun_out
cap() { local name=$1 pat=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$pat -s 6 -c 1 -f -o /tmp/$name "$@" > gpurun_out/r02x_$name.log 2>&1
  echo "exit $? ($name)" | tee -a gpurun_out/r02x_summary.log
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r02_ncu_final_$name.raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/r02_ncu_final_$name.source.csv 2>/dev/null
}
cap k_spmv4 k_spmv4 python scripts/prof_run.py
cap k_update2 k_update2 python scripts/prof_run.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_256_final.csv python bench.py --steps 1 --warmup 1 --iters 40 --no-cpu-baseline --no-parity > gpurun_out/r02x_launches.log 2>&1
echo "exit $? (launches)" | tee -a gpurun_out/r02x_summary.log
