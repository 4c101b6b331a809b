#!/bin/bash
# 2-GPU call: correctness of the boundary-first update order, and device timelines of the 8-GPU-sized slab (256x256x32 per GPU)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02e_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02e_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02e_summary.log
  grep -v "^W1\|^\*\*\*\|OMP_NUM" "gpurun_out/r02e_$name.log" | tail -n 14 | cut -c1-300 | tee -a gpurun_out/r02e_summary.log
}
nvidia-smi --query-gpu=index,name --format=csv | tee -a gpurun_out/r02e_summary.log
run mgpu_check 300 $TR tests/mgpu_check.py p2p+store
run mgpu_check_t41 300 env B200LS_TILE=41 $TR tests/mgpu_check.py p2p+store
run mgpu_check_t53 300 env B200LS_TILE=53 $TR tests/mgpu_check.py p2p+store
run trace1_slab_default 120 python scripts/trace_run.py --size 256 256 32
run trace1_slab_t41 120 python scripts/trace_run.py --size 256 256 32 --tune tile=41
run trace1_slab_t53 120 python scripts/trace_run.py --size 256 256 32 --tune tile=53
run trace2_slab_default 120 $TR scripts/trace_run.py --size 256 256 64
run trace2_slab_t41 120 $TR scripts/trace_run.py --size 256 256 64 --tune tile=41
run trace2_slab_t53 120 $TR scripts/trace_run.py --size 256 256 64 --tune tile=53
run trace2_slab_t18 120 $TR scripts/trace_run.py --size 256 256 64 --tune tile=18
run bench2 200 $TR bench.py --gpus 2 --steps 3
