#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layer.py tests/test_gpu_stages.py -q -x --timeout 200 > gpurun_out/pytest_d.log 2>&1; echo "pytest exit=$?"; tail -n 3 gpurun_out/pytest_d.log
for w in 0 1; do
  B200MOE_WARM=$w timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3_warm$w.log 2>&1
  B200MOE_WARM=$w timeout 300 python bench.py --steps 200 --warmup 10 --workload cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1_warm$w.log 2>&1
done
python tools/bench_summary.py gpurun_out/bench_cfg3_warm*.log gpurun_out/bench_cfg1_warm*.log
python tools/ffn_trace.py 3200 1 > gpurun_out/trace_3200_warm.txt 2>&1; sed -n 1,50p gpurun_out/trace_3200_warm.txt
