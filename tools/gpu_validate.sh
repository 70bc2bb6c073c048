#!/bin/bash
# Final validation of the shipped library: parity suite, smoke, bench lines of the workloads the last changes touch.
mkdir -p gpurun_out
O=gpurun_out
: > $O/final_summary.txt
run() { local name=$1; local to=$2; shift 2; timeout "$to" "$@" > "$O/final_$name.log" 2>&1; echo "$name exit=$?" | tee -a $O/final_summary.txt; }
run pytest 1200 python -m pytest tests -m gpu -q --timeout 300
run smoke  200 python __graft_entry__.py smoke
run bench_cfg3 900 python bench.py --steps 200 --warmup 10 --sustain 3 --encoder
run bench_cfg2 400 python bench.py --steps 300 --warmup 10 --workload cfg2 --no-cpu-baseline --encoder
run bench_cfg1 300 python bench.py --steps 300 --warmup 10 --workload cfg1 --no-cpu-baseline
run bench_cfg3f 300 python bench.py --steps 100 --warmup 10 --workload cfg3f --no-cpu-baseline
run bench_big 400 python bench.py --steps 10 --warmup 3 --workload big --no-cpu-baseline
run bench_cfg4 400 python bench.py --steps 20 --warmup 3 --workload cfg4 --no-cpu-baseline
run bench_sweep 900 python bench.py --workload sweep --steps 20 --warmup 3
tail -n 3 $O/final_pytest.log $O/final_smoke.log
python tools/bench_summary.py $O/final_bench_cfg3.log $O/final_bench_cfg2.log $O/final_bench_cfg1.log $O/final_bench_cfg3f.log $O/final_bench_big.log $O/final_bench_cfg4.log | cut -c1-330
python tools/bench_summary.py $O/final_bench_sweep.log | cut -c1-200
cat $O/final_summary.txt
