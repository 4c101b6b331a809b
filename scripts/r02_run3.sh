#!/bin/bash
# ncu --set full captures; the reports stay on the box (too large), their raw and source pages come back as CSV
set -u
mkdir -p gpurun_out
cap() { local name=$1 pat=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02c_summary.log
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$pat -s 6 -c 1 -f -o /tmp/$name "$@" > gpurun_out/r02c_$name.log 2>&1
  echo "exit $? ($name)" | tee -a gpurun_out/r02c_summary.log
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$name.raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/r02_ncu_$name.source.csv 2>/dev/null
  ls -la gpurun_out/r02_ncu_$name.*
}
cap k_spmv4_t41_kz43 k_spmv4 python scripts/prof_run.py --tune tile=41 kz_chunk=43
cap k_spmv4_t41_kz256 k_spmv4 python scripts/prof_run.py --tune tile=41 kz_chunk=256
cap k_spmv2_t18 k_spmv2 python scripts/prof_run.py --tune tile=18
cap k_update2 k_update2 python scripts/prof_run.py
