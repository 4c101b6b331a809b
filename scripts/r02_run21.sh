#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534"
run() { local name=$1; shift
  echo "=== $name" | tee -a gpurun_out/r02u_summary.log
  timeout 120 "$@" > "gpurun_out/r02u_$name.log" 2>&1
  grep "k_update\|us/iter" "gpurun_out/r02u_$name.log" | grep -v Guessing | cut -c1-180 | tee -a gpurun_out/r02u_summary.log
}
for rep in 1 2; do
run extra1_items2_$rep env B200LS_PUSH_EXTRA=1 B200LS_PUSH_ITEMS=2 $TR scripts/trace_run.py --size 256 256 256
run extra0_items2_$rep env B200LS_PUSH_EXTRA=0 B200LS_PUSH_ITEMS=2 $TR scripts/trace_run.py --size 256 256 256
run extra0_items8_$rep env B200LS_PUSH_EXTRA=0 B200LS_PUSH_ITEMS=8 $TR scripts/trace_run.py --size 256 256 256
run extra1_items8_$rep env B200LS_PUSH_EXTRA=1 B200LS_PUSH_ITEMS=8 $TR scripts/trace_run.py --size 256 256 256
done
