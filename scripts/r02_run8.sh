#!/bin/bash
# 4-GPU call: BASELINE config 4 (256 x 128 x 128 stretched, domain-decomposed), DMDA boxes, 2-D, replicated general systems
set -u
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02h_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02h_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02h_summary.log
  grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func\|ProcessGroupNCCL\|NCCL version" "gpurun_out/r02h_$name.log" | tail -n 30 | cut -c1-400 | tee -a gpurun_out/r02h_summary.log
}
run c4_${N}gpu 400 $TR tests/mgpu_check.py --c4 p2p+store
run mgpu_new_${N}gpu 400 env B200_MGPU_BOX=1 $TR tests/mgpu_check.py p2p+store
run trace_256_${N}gpu 120 $TR scripts/trace_run.py --size 256 256 256
run bench_${N}gpu 300 $TR bench.py --gpus $N --steps 3
