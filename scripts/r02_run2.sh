#!/bin/bash
# Round 2, second GPU call: the TMA kernel (k_spmv4) and the blocked LU.  Every step under its own timeout.
set -u
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02b_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02b_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02b_summary.log
  tail -n 12 "gpurun_out/r02b_$name.log" | cut -c1-300 | tee -a gpurun_out/r02b_summary.log
}
run parity_tile40 200 env B200LS_TILE=40 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run parity_tile41 200 env B200LS_TILE=41 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run parity_tile42 200 env B200LS_TILE=42 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run parity_tile43 200 env B200LS_TILE=43 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run tune_256 200 python scripts/tune_k1.py --tiles 18 40 41 42 43
run tune_256_kz 300 python scripts/tune_k1.py --tiles 40 41 --kz 16 26 32 43 64 128 256
run tune_slab 200 python scripts/tune_k1.py --size 256 256 32 --tiles 18 40 41 42 43 --kz 0 8 11 16 32
run tune_512 200 python scripts/tune_k1.py --size 512 512 512 --tiles 18 40 41 --reps 5
run direct 200 python -m pytest tests/test_zzz_gpu_3_direct.py -m gpu -q
run dense_bench 300 python scripts/dense_bench.py
run bench_40 200 env B200LS_TILE=40 python bench.py --no-cpu-baseline
run bench_41 200 env B200LS_TILE=41 python bench.py --no-cpu-baseline
