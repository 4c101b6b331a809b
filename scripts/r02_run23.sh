#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1; shift
  timeout 250 "$@" > "gpurun_out/r02w_$name.log" 2>&1
  echo "$name: exit $? $(grep -o '"value": [0-9.]*' gpurun_out/r02w_$name.log | head -1) $(grep -o '"avg_launch_us": [0-9.]*' gpurun_out/r02w_$name.log | head -1) $(tail -1 gpurun_out/r02w_$name.log | cut -c1-60)" | tee -a gpurun_out/r02w_summary.log
}
B="python bench.py --no-cpu-baseline --no-parity"
run parity_t48 env B200LS_TILE=48 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run default $B
run t48_kz64 $B --tune tile=48 kz_chunk=64
run t48_kz86 $B --tune tile=48 kz_chunk=86
run t48_kz43 $B --tune tile=48 kz_chunk=43
run t49_kz64 $B --tune tile=49 kz_chunk=64
run default_b $B
