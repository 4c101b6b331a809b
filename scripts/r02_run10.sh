#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02j_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02j_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02j_summary.log
  tail -n 12 "gpurun_out/r02j_$name.log" | cut -c1-700 | tee -a gpurun_out/r02j_summary.log
}
run pytest_gpu 1200 python -m pytest tests -m gpu -q -x
run velocity_128 300 python scripts/velocity_bench.py --size 128 128 128
run velocity_128_graph 300 python scripts/velocity_bench.py --size 128 128 128 --graph 1 --no-cpu
run velocity_2d 300 python scripts/velocity_bench.py --size 448 448 --graph 1 --iters 40
run step_128 300 python scripts/step_bench.py --size 128 128 128
run step_256 300 python scripts/step_bench.py --size 256 256 256
run step_256_jacobi 300 python scripts/step_bench.py --size 256 256 256 --pc jacobi --steps 2
run bench 300 python bench.py
