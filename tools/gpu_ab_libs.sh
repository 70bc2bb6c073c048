#!/bin/bash
# A/B/C.. of several builds of the library on ONE box: every tools/bin/ab_*.so in turn, order reversed in the second round.
#   usage: tools/gpu_ab_libs.sh [workloads...]
mkdir -p gpurun_out
L=3m-asr-inference_b200/libb200moe.so
cp $L /tmp/keep.so
WLS=${@:-cfg3 cfg1}
LIBS=$(ls tools/bin/ab_*.so)
RLIBS=$(ls -r tools/bin/ab_*.so)
export B200MOE_AB_OLD_LIB=1
for round in 1 2; do
  [ $round == 2 ] && LIBS=$RLIBS
  for lib in $LIBS; do
    v=$(basename $lib .so)
    cp $lib $L
    for wl in $WLS; do
      steps=200; [ $wl == big ] && steps=10
      timeout 300 python bench.py --steps $steps --warmup 10 --workload $wl --no-cpu-baseline > gpurun_out/abx_${wl}_${v}_$round.log 2>&1
    done
  done
done
cp /tmp/keep.so $L
for wl in $WLS; do python tools/bench_summary.py gpurun_out/abx_${wl}_*.log | cut -c1-130; done
