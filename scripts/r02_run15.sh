#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02o_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02o_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02o_summary.log
  tail -n 6 "gpurun_out/r02o_$name.log" | cut -c1-300 | tee -a gpurun_out/r02o_summary.log
}
run pytest_gpu 1200 python -m pytest tests -m gpu -q -x
run bench_default 200 python bench.py --no-cpu-baseline
run bench_t18 200 python bench.py --no-cpu-baseline --no-parity --tune tile=18
run bench_512 200 python bench.py --no-cpu-baseline --no-parity --size 512 512 512 --iters 200 --steps 3
run bench_512_t18 200 python bench.py --no-cpu-baseline --no-parity --size 512 512 512 --iters 200 --steps 3 --tune tile=18
run bench_jacobi 200 python bench.py --no-cpu-baseline --no-parity --pc jacobi
run bench_jacobi_t18 200 python bench.py --no-cpu-baseline --no-parity --pc jacobi --tune tile=18
run bench_384 200 python bench.py --no-cpu-baseline --no-parity --size 384 320 200 --iters 200 --steps 3
run bench_384_t18 200 python bench.py --no-cpu-baseline --no-parity --size 384 320 200 --iters 200 --steps 3 --tune tile=18
