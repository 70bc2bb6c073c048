#!/bin/bash
# round 2, session 2: staged-operator tests, memory-system facts, baseline cfg3 line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fmoe_functions.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_functions.log 2>&1; echo "pytest exit=$?"
tail -n 15 gpurun_out/pytest_functions.log
timeout 300 tools/bin/mem_facts_bench > gpurun_out/mem_facts.txt 2>&1; echo "mem_facts exit=$?"
cat gpurun_out/mem_facts.txt
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3_s2base.log 2>&1
python tools/bench_summary.py gpurun_out/bench_cfg3_s2base.log
