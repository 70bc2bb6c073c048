#!/bin/bash
mkdir -p gpurun_out
for w in 0 1 2; do
  B200MOE_WPOL=$w timeout 300 python bench.py --steps 200 --warmup 20 --workload cfg3x1 --no-cpu-baseline > gpurun_out/bench_cfg3x1_wpol$w.log 2>&1
done
python tools/bench_summary.py gpurun_out/bench_cfg3x1_wpol*.log
