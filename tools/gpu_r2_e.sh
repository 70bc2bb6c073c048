#!/bin/bash
mkdir -p gpurun_out
for w in 0 4 8 16 32; do
  B200MOE_WARM=$w timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3_warm$w.log 2>&1
  B200MOE_WARM=$w python tools/ffn_trace.py 3200 1 > gpurun_out/trace_3200_warm$w.txt 2>&1; echo "warm $w"; sed -n 1,2p gpurun_out/trace_3200_warm$w.txt; grep "mma_stage_issued\|mma_issued  \|mma_first_data" gpurun_out/trace_3200_warm$w.txt | head -8
done
python tools/bench_summary.py gpurun_out/bench_cfg3_warm*.log
