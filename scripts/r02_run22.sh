#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1; shift
  timeout 250 "$@" > "gpurun_out/r02v_$name.log" 2>&1
  echo "$name: exit $? $(grep -o '"value": [0-9.]*' gpurun_out/r02v_$name.log | head -1) $(grep -o '"avg_launch_us": [0-9.]*' gpurun_out/r02v_$name.log | head -1) $(tail -1 gpurun_out/r02v_$name.log | cut -c1-80)" | tee -a gpurun_out/r02v_summary.log
}
B="python bench.py --no-cpu-baseline --no-parity"
run parity_t46 env B200LS_TILE=46 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x
run parity_t47 env B200LS_TILE=47 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run default $B
run t46 $B --tune tile=46
run t47 $B --tune tile=47
run default_b $B
run t46_b $B --tune tile=46
run default_512 $B --size 512 512 512 --iters 200 --steps 3
run t46_512 $B --size 512 512 512 --iters 200 --steps 3 --tune tile=46
run t46_slab python scripts/trace_run.py --size 256 256 32 --tune tile=46
run t41_slab python scripts/trace_run.py --size 256 256 32
