#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1; shift
  timeout 200 "$@" > "gpurun_out/r02y_$name.log" 2>&1
  echo "$name: exit $? $(grep -o '"value": [0-9.]*' gpurun_out/r02y_$name.log | head -1) $(grep -o '"avg_launch_us": [0-9.]*' gpurun_out/r02y_$name.log | head -1) $(tail -1 gpurun_out/r02y_$name.log | cut -c1-50)" | tee -a gpurun_out/r02y_summary.log
}
B="python bench.py --no-cpu-baseline --no-parity"
run parity_t62 env B200LS_TILE=62 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run default $B
run t62_kz64 $B --tune tile=62 kz_chunk=64
run t62_kz43 $B --tune tile=62 kz_chunk=43
run t60_kz52 $B --tune tile=60 kz_chunk=52
run default_b $B
