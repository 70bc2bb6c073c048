#!/bin/bash
# A/B of two builds of the library on ONE box: tools/bin/libb200moe_old.so against the in-tree one
mkdir -p gpurun_out
L=3m-asr-inference_b200/libb200moe.so
cp $L /tmp/new.so
for round in 1 2; do
for v in new old; do
  if [ $v == old ]; then cp tools/bin/libb200moe_old.so $L; else cp /tmp/new.so $L; fi
  timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3_ab_${v}_$round.log 2>&1
  timeout 300 python bench.py --steps 300 --warmup 10 --workload cfg2 --no-cpu-baseline > gpurun_out/bench_cfg2_ab_${v}_$round.log 2>&1
done
done
cp /tmp/new.so $L
python tools/bench_summary.py gpurun_out/bench_cfg3_ab_*.log gpurun_out/bench_cfg2_ab_*.log | cut -c1-120
timeout 300 python -m pytest tests/test_gpu_layer.py tests/test_gpu_block.py -q -x --timeout 300 2>&1 | tail -2
