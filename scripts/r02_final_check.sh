#!/bin/bash
# Final validation of the round (2-GPU box): full device suite incl. the 2-GPU test, multi-GPU checks, smoke, both bench arms
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02z_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02z_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02z_summary.log
  grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func\|ProcessGroupNCCL\|NCCL version" "gpurun_out/r02z_$name.log" | tail -n 6 | cut -c1-400 | tee -a gpurun_out/r02z_summary.log
}
run pytest_gpu 1200 python -m pytest tests -m gpu -q -x
run smoke 300 python __graft_entry__.py --smoke
run mgpu_check 400 env B200_MGPU_BOX=1 $TR tests/mgpu_check.py p2p+store
run bench 300 python bench.py
run bench_reference_short 300 python bench.py --impl reference --steps 1 --warmup 1 --iters 100
run bench2 300 $TR bench.py --gpus 2
