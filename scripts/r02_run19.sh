#!/bin/bash
# 8-GPU call (final): scaling of the headline on one box, device timeline, config 5 (512^3) at N = 1 / 8
set -u
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534"
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535"
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02s_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02s_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02s_summary.log
  grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func\|ProcessGroupNCCL\|NCCL version" "gpurun_out/r02s_$name.log" | tail -n 8 | cut -c1-600 | tee -a gpurun_out/r02s_summary.log
}
run bench_8gpu 300 $TR8 bench.py --gpus 8 --steps 10 --warmup 3
run trace_256_8gpu 120 $TR8 scripts/trace_run.py --size 256 256 256
run bench_4gpu 300 $TR4 bench.py --gpus 4 --steps 10 --warmup 3
run bench_2gpu 300 $TR2 bench.py --gpus 2 --steps 10 --warmup 3
run bench_1gpu 200 python bench.py --no-cpu-baseline --steps 10
run bench_512_8gpu 300 $TR8 bench.py --gpus 8 --steps 3 --size 512 512 512 --iters 200 --no-parity
run bench_512_1gpu 300 python bench.py --steps 3 --size 512 512 512 --iters 200 --no-parity --no-cpu-baseline
run c4_4gpu 300 $TR4 tests/mgpu_check.py --c4 p2p+store
