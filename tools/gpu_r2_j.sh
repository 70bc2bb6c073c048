#!/bin/bash
mkdir -p gpurun_out
for v in "0 4 0" "1 4 0" "1 5 0" "1 5 1"; do
  set -- $v
  B200MOE_PREFETCH=$1 B200MOE_PDL_TRIG=$2 B200MOE_FFN_PREFETCH=$3 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3_l2pf$1_trig$2_pf$3.log 2>&1
  B200MOE_PREFETCH=$1 B200MOE_PDL_TRIG=$2 B200MOE_FFN_PREFETCH=$3 timeout 300 python bench.py --steps 200 --warmup 10 --workload cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1_l2pf$1_trig$2_pf$3.log 2>&1
done
python tools/bench_summary.py gpurun_out/bench_cfg3_l2pf*.log gpurun_out/bench_cfg1_l2pf*.log
B200MOE_PREFETCH=1 B200MOE_FFN_PREFETCH=0 timeout 200 python tools/timeline.py 3200 6 > gpurun_out/timeline_l2pf1.txt 2>&1; cat gpurun_out/timeline_l2pf1.txt
