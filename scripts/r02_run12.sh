#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
run() { local name=$1; shift
  echo "=== $name" | tee -a gpurun_out/r02l_summary.log
  timeout 120 "$@" > "gpurun_out/r02l_$name.log" 2>&1
  grep "k_update\|us/iter" "gpurun_out/r02l_$name.log" | cut -c1-200 | tee -a gpurun_out/r02l_summary.log
}
for rep in 1 2; do
run flags_$rep $TR scripts/trace_run.py --size 256 256 64
run noflags_$rep env B200LS_NO_HALO_FLAGS=1 $TR scripts/trace_run.py --size 256 256 64
run nofence_$rep env B200LS_NO_HALO_FLAGS=1 B200LS_DBG_FLAGS=1 $TR scripts/trace_run.py --size 256 256 64
run flags_items1_$rep env B200LS_PUSH_ITEMS=1 $TR scripts/trace_run.py --size 256 256 64
done
