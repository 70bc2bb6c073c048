#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3_span.log 2>&1; echo "exit=$?"
timeout 300 python bench.py --steps 200 --warmup 10 --workload cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1_span.log 2>&1; echo "exit=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --workload big --no-cpu-baseline > gpurun_out/bench_big_span.log 2>&1; echo "exit=$?"
python tools/bench_summary.py gpurun_out/bench_cfg3_span.log gpurun_out/bench_cfg1_span.log gpurun_out/bench_big_span.log | cut -c1-330
tail -3 gpurun_out/bench_cfg3_span.log | grep -v "^{" | tail -3
