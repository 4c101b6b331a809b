#!/bin/bash
# 8-GPU call: scaling of the headline (N = 8 and N = 1 on the same box), device timeline, DMDA 2x2x2 boxes
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02i_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02i_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02i_summary.log
  grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func\|ProcessGroupNCCL\|NCCL version" "gpurun_out/r02i_$name.log" | tail -n 30 | cut -c1-600 | tee -a gpurun_out/r02i_summary.log
}
run bench_8gpu 300 $TR bench.py --gpus 8 --steps 5
run trace_256_8gpu 120 $TR scripts/trace_run.py --size 256 256 256
run trace_256_8gpu_t18 120 $TR scripts/trace_run.py --size 256 256 256 --tune tile=18
run trace_256_8gpu_t53 120 $TR scripts/trace_run.py --size 256 256 256 --tune tile=53
run mgpu_new_8gpu 400 env B200_MGPU_BOX=1 $TR tests/mgpu_check.py p2p+store
run bench_1gpu 200 python bench.py --no-cpu-baseline
run bench_512_8gpu 300 $TR bench.py --gpus 8 --steps 3 --size 512 512 512 --iters 200 --no-parity
