#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02r_summary.log
  timeout "$t" "$@" > "gpurun_out/r02r_$name.log" 2>&1
  echo "exit $? ($name)" | tee -a gpurun_out/r02r_summary.log
  grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func\|ProcessGroupNCCL\|NCCL version" "gpurun_out/r02r_$name.log" | tail -n 22 | cut -c1-330 | tee -a gpurun_out/r02r_summary.log
}
run mgpu_check 400 env B200_MGPU_BOX=1 $TR tests/mgpu_check.py p2p+store
run trace2_256 120 $TR scripts/trace_run.py --size 256 256 256
run trace2_slab 120 $TR scripts/trace_run.py --size 256 256 64
run trace2_256_b 120 $TR scripts/trace_run.py --size 256 256 256
run bench2 200 $TR bench.py --gpus 2 --steps 5
