#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layer.py tests/test_gpu_stages.py tests/test_gpu_ep.py -q -x --timeout 200 > gpurun_out/pytest_c.log 2>&1; echo "pytest exit=$?"; tail -n 3 gpurun_out/pytest_c.log
timeout 300 python bench.py --steps 50 --warmup 5 --workload cfg4 --no-cpu-baseline > gpurun_out/bench_cfg4_c.log 2>&1
python tools/bench_summary.py gpurun_out/bench_cfg4_c.log
timeout 300 tools/bin/stream_bench > gpurun_out/stream_bench.txt 2>&1; cat gpurun_out/stream_bench.txt
