#!/bin/bash
mkdir -p gpurun_out
B200MOE_FIT_B=1 timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_layer.py tests/test_gpu_block.py tests/test_gpu_ep.py -q -x --timeout 300 2>&1 | tail -2
for round in 1 2; do
for v in 0 1; do
  for wl in cfg3 cfg2 cfg4; do
    B200MOE_FIT_B=$v timeout 300 python bench.py --steps 200 --warmup 10 --workload $wl --no-cpu-baseline > gpurun_out/fitb_${wl}_v${v}_$round.log 2>&1
  done
done
done
for wl in cfg3 cfg2 cfg4; do python tools/bench_summary.py gpurun_out/fitb_${wl}_*.log | cut -c1-190; done
