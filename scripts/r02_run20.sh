#!/bin/bash
set -u
mkdir -p gpurun_out
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534"
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535"
run() { local name=$1 t=$2; shift 2
  timeout "$t" "$@" > "gpurun_out/r02t_$name.log" 2>&1
  echo "exit $? ($name)" | tee -a gpurun_out/r02t_summary.log
}
run bench_4gpu 300 $TR4 bench.py --gpus 4 --steps 10 --warmup 3
run bench_2gpu 300 $TR2 bench.py --gpus 2 --steps 10 --warmup 3
run bench_1gpu 200 python bench.py --no-cpu-baseline --steps 10
run trace_4gpu 120 $TR4 scripts/trace_run.py --size 256 256 256
