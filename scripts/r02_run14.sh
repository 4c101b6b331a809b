#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02n_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02n_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02n_summary.log
  tail -n 6 "gpurun_out/r02n_$name.log" | cut -c1-420 | tee -a gpurun_out/r02n_summary.log
}
run direct 200 python -m pytest tests/test_zzz_gpu_3_direct.py -m gpu -q
run dense_bench 300 python scripts/dense_bench.py 500 2000 4096 8192
run dense_bench_one_cta 300 env B200LS_DENSE_ONE_CTA=1 python scripts/dense_bench.py 2000
run bench_default 200 python bench.py --no-cpu-baseline --no-parity
run bench_t41_kz43 200 python bench.py --no-cpu-baseline --no-parity --tune tile=41 kz_chunk=43
run bench_t41_kz32 200 python bench.py --no-cpu-baseline --no-parity --tune tile=41 kz_chunk=32
run bench_t40_kz43 200 python bench.py --no-cpu-baseline --no-parity --tune tile=40 kz_chunk=43
run bench_default2 200 python bench.py --no-cpu-baseline --no-parity
