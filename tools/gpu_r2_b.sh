#!/bin/bash
# Round-2 GPU pass B: route 2-tile fix check + FFN experiments (what bounds the weight-streaming regime?)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layer.py tests/test_gpu_stages.py -q -x --timeout 200 > gpurun_out/pytest_b.log 2>&1; echo "pytest exit=$?"; tail -n 3 gpurun_out/pytest_b.log
timeout 300 python bench.py --steps 50 --warmup 5 --workload cfg4 --no-cpu-baseline > gpurun_out/bench_cfg4_b.log 2>&1
for d in 0 8 16 24 32 56 57 59; do
  B200MOE_DBG=$d timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3_dbg$d.log 2>&1
done
for d in 0 24 56; do
  B200MOE_DBG=$d timeout 200 python tools/ffn_trace.py 3200 2 > gpurun_out/trace_3200_dbg$d.txt 2>&1
done
python tools/bench_summary.py gpurun_out/bench_cfg4_b.log gpurun_out/bench_cfg3_dbg*.log
head -18 gpurun_out/trace_3200_dbg0.txt gpurun_out/trace_3200_dbg24.txt gpurun_out/trace_3200_dbg56.txt
