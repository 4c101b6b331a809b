#!/bin/bash
# tiled line-coefficient kernels with the per-thread cp.async queue: device tests, then the velocity BiCGStab per variant
set -u
mkdir -p gpurun_out
run() { local name=$1; shift
  timeout 100 "$@" > "gpurun_out/r02C_$name.log" 2>&1
  echo "$name: exit $? $(tail -1 gpurun_out/r02C_$name.log | cut -c1-120)" | tee -a gpurun_out/r02C_summary.log
}
run pytest_staggered python -m pytest tests/test_zzz_gpu_2_staggered.py -m gpu -q -x
run vel128 python scripts/velocity_bench.py --size 128 128 128 --no-cpu --no-csr --reps 2 --tiles 0 2 --zchunks 0 8 16 32 --stages 3 4
grep -h iterations_per_s gpurun_out/r02C_vel*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['size'], 'tile', d['sep_tile'], 'zchunk', d['sep_zchunk'], 'stages', d['sep_stages'], d['iterations_per_s'], 'it/s', d['model_GBs'], 'GB/s')
" | tee -a gpurun_out/r02C_summary.log
