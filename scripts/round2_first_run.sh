#!/bin/bash
# First GPU call of the next round (one B200): everything that was written after round 1's GPU budget ran out is run
# and timed here, cheapest and most informative first, each step under its own timeout so that a hang costs minutes,
# not the call.  Usage:
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/round2_first_run.sh'
# Outputs land in gpurun_out/ (copy what matters to profiles/r02_*).
set -u
mkdir -p gpurun_out
run() { # name timeout command...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02_summary.log
  timeout "$t" "$@" > "gpurun_out/r02_$name.log" 2>&1
  echo "exit $? ($name)" | tee -a gpurun_out/r02_summary.log
  tail -n 6 "gpurun_out/r02_$name.log" | tee -a gpurun_out/r02_summary.log
}
# 1. the validated suite first (must stay green), then the new device tests one file at a time
run pytest_validated 900 python -m pytest tests -m gpu -q -x --deselect tests/test_zzz_gpu_1_boxes.py --deselect tests/test_zzz_gpu_2_staggered.py --deselect tests/test_zzz_gpu_4_multigrid.py --deselect tests/test_zzz_gpu_3_ops.py --deselect tests/test_zzz_gpu_3_direct.py
run pytest_shim_new 300 python -m pytest tests/test_gpu_shim.py -m gpu -q --runxfail
run pytest_boxes 300 python -m pytest tests/test_zzz_gpu_1_boxes.py -m gpu -q
run pytest_staggered 300 python -m pytest tests/test_zzz_gpu_2_staggered.py -m gpu -q --runxfail
run pytest_ops 300 python -m pytest tests/test_zzz_gpu_3_ops.py -m gpu -q --runxfail
run pytest_direct 300 python -m pytest tests/test_zzz_gpu_3_direct.py -m gpu -q --runxfail
run pytest_multigrid 600 python -m pytest tests/test_zzz_gpu_4_multigrid.py -m gpu -q --runxfail
# the optional multigrid paths through the same tests (environment switches read by b200ls_create)
run pytest_multigrid_tail 600 env B200LS_MG_TAIL=1 python -m pytest tests/test_zzz_gpu_4_multigrid.py -m gpu -q --runxfail
run pytest_multigrid_fuse 600 env B200LS_MG_FUSE=1 python -m pytest tests/test_zzz_gpu_4_multigrid.py -m gpu -q --runxfail
run pytest_multigrid_graph 600 env B200LS_MG_GRAPH=1 B200LS_MG_TAIL=1 B200LS_MG_FUSE=1 python -m pytest tests/test_zzz_gpu_4_multigrid.py -m gpu -q --runxfail
run pytest_csr_graph 600 env B200LS_CSR_GRAPH=1 python -m pytest tests/test_gpu_csr.py tests/test_velocity_operator.py tests/test_zzz_gpu_2_staggered.py -m gpu -q --runxfail
# 2. round-2 kernel candidates against the default: 256^3 and the 8-GPU slab
run tune_256 600 python scripts/tune_k1.py --tiles 10 18 30 31 32 33
run tune_slab 300 python scripts/tune_k1.py --size 256 256 32 --tiles 10 13 18 30 31 32 33
run parity_tile30 300 env B200LS_TILE=30 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run parity_tile32 300 env B200LS_TILE=32 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run parity_jacobi_fly 300 env B200LS_UPD_VARIANT=2 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
run bench_jacobi 300 python bench.py --pc jacobi --no-cpu-baseline
run bench_jacobi_fly 300 env B200LS_UPD_VARIANT=2 python bench.py --pc jacobi --no-cpu-baseline
# 3. time to solution: none / jacobi / mg
run tts_128 300 python scripts/tts_bench.py --size 128 128 128
run tts_256 600 python scripts/tts_bench.py --size 256 256 256 --pcs jacobi mg
run tts_128_graph 300 python scripts/tts_bench.py --size 128 128 128 --pcs mg --mg-graph 1
run tts_128_tail 300 python scripts/tts_bench.py --size 128 128 128 --pcs mg --mg-tail 1
run tts_256_tail 300 python scripts/tts_bench.py --size 256 256 256 --pcs mg --mg-tail 1
run tts_256_tail_fuse 300 python scripts/tts_bench.py --size 256 256 256 --pcs mg --mg-tail 1 --mg-fuse 1
run tts_2d_tail_graph 300 python scripts/tts_bench.py --size 448 448 --pcs mg --mg-tail 1 --mg-graph 1
run tts_2d_graph 300 python scripts/tts_bench.py --size 448 448 --pcs mg --mg-graph 1
run tts_2d 300 python scripts/tts_bench.py --size 448 448 --pcs jacobi mg
run tts_256_stretched 600 python scripts/tts_bench.py --size 256 256 256 --pcs mg --stretched
# 4. the headline bench and the launch list of the same command
run bench 600 python bench.py
run launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_256.csv python bench.py --steps 1 --warmup 1 --iters 40 --no-cpu-baseline
# 5. on a box with several GPUs (gpurun --gpus 2): the multi-GPU paths that have not run on GPUs yet
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'B200_MGPU_BOX=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py p2p+store > gpurun_out/r02_mgpu_new.log 2>&1; tail -30 gpurun_out/r02_mgpu_new.log'
