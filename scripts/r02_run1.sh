#!/bin/bash
# Round 2, first GPU call (one B200): run + time everything written blind in round 1.  Outputs: gpurun_out/r02_*.
set -u
mkdir -p gpurun_out
run() { # name timeout command...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02_summary.log
  tail -n 8 "gpurun_out/r02_$name.log" | cut -c1-400 | tee -a gpurun_out/r02_summary.log
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv | tee -a gpurun_out/r02_summary.log
run pytest_gpu 1200 python -m pytest tests -m gpu -q -x
run pytest_multigrid_tail 300 env B200LS_MG_TAIL=1 python -m pytest tests/test_zzz_gpu_4_multigrid.py -m gpu -q
run pytest_multigrid_fuse 300 env B200LS_MG_FUSE=1 python -m pytest tests/test_zzz_gpu_4_multigrid.py -m gpu -q
run pytest_multigrid_graph 300 env B200LS_MG_GRAPH=1 B200LS_MG_TAIL=1 B200LS_MG_FUSE=1 python -m pytest tests/test_zzz_gpu_4_multigrid.py -m gpu -q
run pytest_csr_graph 300 env B200LS_CSR_GRAPH=1 python -m pytest tests/test_gpu_csr.py tests/test_velocity_operator.py tests/test_zzz_gpu_2_staggered.py -m gpu -q
run tune_256 300 python scripts/tune_k1.py --tiles 10 18 30 31 32 33
run tune_slab 300 python scripts/tune_k1.py --size 256 256 32 --tiles 10 13 18 30 31 32 33
run parity_tile30 300 env B200LS_TILE=30 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run parity_tile32 300 env B200LS_TILE=32 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run parity_jacobi_fly 300 env B200LS_UPD_VARIANT=2 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run bench_jacobi 200 python bench.py --pc jacobi --no-cpu-baseline
run bench_jacobi_fly 200 env B200LS_UPD_VARIANT=2 python bench.py --pc jacobi --no-cpu-baseline
run tts_128 200 python scripts/tts_bench.py --size 128 128 128
run tts_256 400 python scripts/tts_bench.py --size 256 256 256 --pcs jacobi mg
run tts_128_graph 200 python scripts/tts_bench.py --size 128 128 128 --pcs mg --mg-graph 1
run tts_128_tail 200 python scripts/tts_bench.py --size 128 128 128 --pcs mg --mg-tail 1
run tts_256_tail 200 python scripts/tts_bench.py --size 256 256 256 --pcs mg --mg-tail 1
run tts_256_tail_fuse 200 python scripts/tts_bench.py --size 256 256 256 --pcs mg --mg-tail 1 --mg-fuse 1
run tts_256_tail_fuse_graph 200 python scripts/tts_bench.py --size 256 256 256 --pcs mg --mg-tail 1 --mg-fuse 1 --mg-graph 1
run tts_2d_tail_graph 200 python scripts/tts_bench.py --size 448 448 --pcs mg --mg-tail 1 --mg-graph 1
run tts_2d 200 python scripts/tts_bench.py --size 448 448 --pcs jacobi mg
run tts_256_stretched 300 python scripts/tts_bench.py --size 256 256 256 --pcs mg --stretched --mg-tail 1
run bench 400 python bench.py
run launches 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_256.csv python bench.py --steps 1 --warmup 1 --iters 40 --no-cpu-baseline --no-parity
# launch list of one multigrid-preconditioned solve at 256^3: which level kernels cost what
run launches_mg 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_mg_256.csv python scripts/tts_bench.py --size 256 256 256 --pcs mg --reps 1
