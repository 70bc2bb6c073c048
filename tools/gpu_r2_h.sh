#!/bin/bash
mkdir -p gpurun_out
for v in "4 1" "5 1" "5 0" "4 0" "7 1"; do
  set -- $v
  B200MOE_PDL_TRIG=$1 B200MOE_FFN_PREFETCH=$2 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3_trig$1_pf$2.log 2>&1
  B200MOE_PDL_TRIG=$1 B200MOE_FFN_PREFETCH=$2 timeout 300 python bench.py --steps 200 --warmup 10 --workload cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1_trig$1_pf$2.log 2>&1
done
python tools/bench_summary.py gpurun_out/bench_cfg3_trig*.log gpurun_out/bench_cfg1_trig*.log
B200MOE_PDL_TRIG=5 timeout 600 python -m pytest tests/test_gpu_layer.py tests/test_gpu_block.py -q -x --timeout 300 > gpurun_out/pytest_h.log 2>&1; echo "pytest exit=$?"
tail -n 3 gpurun_out/pytest_h.log
