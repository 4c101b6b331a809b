#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/r02m_summary.log
  local t0=$SECONDS
  timeout "$t" "$@" > "gpurun_out/r02m_$name.log" 2>&1
  echo "exit $? ($name) $((SECONDS-t0)) s" | tee -a gpurun_out/r02m_summary.log
  tail -n 8 "gpurun_out/r02m_$name.log" | cut -c1-700 | tee -a gpurun_out/r02m_summary.log
}
run pytest_gpu 1200 python -m pytest tests -m gpu -q -x
run step_128 300 python scripts/step_bench.py --size 128 128 128
run step_256 300 python scripts/step_bench.py --size 256 256 256
run tts_256 300 python scripts/tts_bench.py --size 256 256 256 --pcs mg --mg-tail 1 --mg-fuse 1 --mg-graph 1
run smoke 300 python __graft_entry__.py --smoke
run launches 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_256.csv python bench.py --steps 1 --warmup 1 --iters 40 --no-cpu-baseline --no-parity
run ncu_traffic 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_spmv2 -s 10 -c 3 --csv --log-file gpurun_out/r02_traffic_k_spmv2.csv python scripts/prof_run.py
