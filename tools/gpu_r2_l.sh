#!/bin/bash
mkdir -p gpurun_out
B200MOE_FFN_PREFETCH=1 B200MOE_PDL_TRIG=5 python tools/timeline.py 3200 5 2>&1 | sed -n "/expert kernel, us/,\$p"
for v in "4 0" "5 1" "4 1"; do
  set -- $v
  B200MOE_PDL_TRIG=$1 B200MOE_FFN_PREFETCH=$2 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3_trig$1_pf$2.log 2>&1
  B200MOE_PDL_TRIG=$1 B200MOE_FFN_PREFETCH=$2 timeout 300 python bench.py --steps 200 --warmup 10 --workload cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1_trig$1_pf$2.log 2>&1
done
python tools/bench_summary.py gpurun_out/bench_cfg3_trig4_pf0.log gpurun_out/bench_cfg3_trig5_pf1.log gpurun_out/bench_cfg3_trig4_pf1.log gpurun_out/bench_cfg1_trig4_pf0.log gpurun_out/bench_cfg1_trig5_pf1.log gpurun_out/bench_cfg1_trig4_pf1.log | cut -c1-150
