#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02d_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02d_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02d_summary.log
  tail -n 14 "gpurun_out/r02d_$name.log" | cut -c1-300 | tee -a gpurun_out/r02d_summary.log
}
run parity_tile50 200 env B200LS_TILE=50 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run parity_tile52 200 env B200LS_TILE=52 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run parity_tile53 200 env B200LS_TILE=53 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run tune_slab 200 python scripts/tune_k1.py --size 256 256 32 --tiles 18 41 43 50 51 52 53 54 55 56
run tune_slab16 200 python scripts/tune_k1.py --size 256 256 16 --tiles 18 41 50 52 54 56
run tune_slab64 200 python scripts/tune_k1.py --size 256 256 64 --tiles 18 41 50 51 53 54 55
run tune_256 200 python scripts/tune_k1.py --tiles 18 41 50 51 53 54 55
run tune_128 200 python scripts/tune_k1.py --size 128 128 128 --tiles 18 41 50 51 53
run bench_41 200 env B200LS_TILE=41 python bench.py --no-cpu-baseline --no-parity
run bench_18 200 python bench.py --no-cpu-baseline --no-parity
