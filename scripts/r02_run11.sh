#!/bin/bash
# 2-GPU call: halo hand-shake flags (pusher fence off the reduction path)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02k_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02k_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02k_summary.log
  grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func\|ProcessGroupNCCL\|NCCL version" "gpurun_out/r02k_$name.log" | tail -n 24 | cut -c1-400 | tee -a gpurun_out/r02k_summary.log
}
run mgpu_check 400 env B200_MGPU_BOX=1 $TR tests/mgpu_check.py p2p+store
run mgpu_check_all 400 $TR tests/mgpu_check.py
run mgpu_check_t18 300 env B200LS_TILE=18 $TR tests/mgpu_check.py p2p+store
run mgpu_check_t53 300 env B200LS_TILE=53 $TR tests/mgpu_check.py p2p+store
run trace2_slab 120 $TR scripts/trace_run.py --size 256 256 64
run trace2_slab_noflags 120 env B200LS_NO_HALO_FLAGS=1 $TR scripts/trace_run.py --size 256 256 64
run trace2_256 120 $TR scripts/trace_run.py --size 256 256 256
run bench2 200 $TR bench.py --gpus 2 --steps 5
run c4 300 $TR tests/mgpu_check.py --c4 p2p+store
