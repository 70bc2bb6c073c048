#!/bin/bash
# A/B of two builds of the library on ONE box (box-to-box spread is +-5 %): tools/bin/libb200moe_old.so against the
# in-tree one, alternating.   usage: tools/gpu_ab_lib.sh [workloads...]
mkdir -p gpurun_out
L=3m-asr-inference_b200/libb200moe.so
cp $L /tmp/new.so
WLS=${@:-cfg3 cfg1}
for round in 1 2; do
for v in new old; do
  if [ $v == old ]; then cp tools/bin/libb200moe_old.so $L; export B200MOE_AB_OLD_LIB=1; else cp /tmp/new.so $L; unset B200MOE_AB_OLD_LIB; fi
  for wl in $WLS; do
    steps=200; [ $wl == big ] && steps=10
    timeout 300 python bench.py --steps $steps --warmup 10 --workload $wl --no-cpu-baseline > gpurun_out/bench_${wl}_ab_${v}_$round.log 2>&1
  done
done
done
unset B200MOE_AB_OLD_LIB
cp /tmp/new.so $L
for wl in $WLS; do python tools/bench_summary.py gpurun_out/bench_${wl}_ab_*.log | cut -c1-150; done
