#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1; shift
  timeout 200 "$@" > "gpurun_out/r02p_$name.log" 2>&1
  echo "$name: $(grep -o '"value": [0-9.]*' gpurun_out/r02p_$name.log | head -1) $(grep -o '"avg_launch_us": [0-9.]*' gpurun_out/r02p_$name.log | head -1)" | tee -a gpurun_out/r02p_summary.log
}
B="python bench.py --no-cpu-baseline --no-parity"
run default $B
run t44_s6 $B --tune tile=44 kz_chunk=43
run t45_s3 $B --tune tile=45 kz_chunk=43
run t41_kz37 $B --tune tile=41 kz_chunk=37
run t41_kz52 $B --tune tile=41 kz_chunk=52
run t41_kz64 $B --tune tile=41 kz_chunk=64
run t41_kz86 $B --tune tile=41 kz_chunk=86
run default_again $B
run upd_blocks_1184 $B --tune upd_blocks=1184
run upd_blocks_296 $B --tune upd_blocks=296
run upd_fwd $B --tune upd_reverse=0
