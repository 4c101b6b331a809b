#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02g_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02g_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02g_summary.log
  grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func\|ProcessGroupNCCL\|NCCL version" "gpurun_out/r02g_$name.log" | tail -n 16 | cut -c1-400 | tee -a gpurun_out/r02g_summary.log
}
run mgpu_check 300 env B200_MGPU_BOX=1 $TR tests/mgpu_check.py p2p+store
run trace2_items8 120 $TR scripts/trace_run.py --size 256 256 64
run trace2_items2 120 env B200LS_PUSH_ITEMS=2 $TR scripts/trace_run.py --size 256 256 64
run trace2_items1 120 env B200LS_PUSH_ITEMS=1 $TR scripts/trace_run.py --size 256 256 64
run trace2_items32 120 env B200LS_PUSH_ITEMS=32 $TR scripts/trace_run.py --size 256 256 64
run trace2_nofence 120 env B200LS_DBG_FLAGS=1 $TR scripts/trace_run.py --size 256 256 64
