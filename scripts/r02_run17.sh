#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02q_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02q_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02q_summary.log
  grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func\|ProcessGroupNCCL\|NCCL version" "gpurun_out/r02q_$name.log" | tail -n 22 | cut -c1-330 | tee -a gpurun_out/r02q_summary.log
}
run pytest_gpu 900 python -m pytest tests -m gpu -q -x
run mgpu_check 400 env B200_MGPU_BOX=1 $TR tests/mgpu_check.py p2p+store
run c4 300 $TR tests/mgpu_check.py --c4 p2p+store
run trace2_256 120 $TR scripts/trace_run.py --size 256 256 256
run trace2_256_t18 120 $TR scripts/trace_run.py --size 256 256 256 --tune tile=18
run bench2 200 $TR bench.py --gpus 2 --steps 5
run bench1 200 python bench.py --no-cpu-baseline
