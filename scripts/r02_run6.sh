#!/bin/bash
# 2-GPU call: dedicated pusher CTAs in k_update2, TMA kernel as the default for cache-resident slabs
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02f_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02f_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02f_summary.log
  grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func\|ProcessGroupNCCL\|NCCL version" "gpurun_out/r02f_$name.log" | tail -n 14 | cut -c1-400 | tee -a gpurun_out/r02f_summary.log
}
run mgpu_check 300 env B200_MGPU_BOX=1 $TR tests/mgpu_check.py p2p+store
run trace2_slab 120 $TR scripts/trace_run.py --size 256 256 64
run trace2_slab_t18 120 $TR scripts/trace_run.py --size 256 256 64 --tune tile=18
run trace2_256 120 $TR scripts/trace_run.py --size 256 256 256
run bench2 200 $TR bench.py --gpus 2 --steps 3
run pytest_two_gpu 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_gpu"
