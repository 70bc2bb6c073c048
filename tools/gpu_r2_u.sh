#!/bin/bash
tools/gpu_ab_lib.sh cfg3 cfg1
for v in "1" "0"; do
B200MOE_FIT_N=$v timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3_fitn$v.log 2>&1
B200MOE_FIT_N=$v timeout 300 python bench.py --steps 200 --warmup 10 --workload cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1_fitn$v.log 2>&1
done
python tools/bench_summary.py gpurun_out/bench_cfg3_fitn*.log gpurun_out/bench_cfg1_fitn*.log | cut -c1-150
