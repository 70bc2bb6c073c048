#!/bin/bash
# Round-2 GPU pass A: full parity suite, smoke, bench on every single-GPU workload, sweep (one GPU).
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1; local to=$2; shift 2; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; }
run pytest 1200 python -m pytest tests -m gpu -q --timeout 300
run smoke  200 python __graft_entry__.py smoke
run bench_cfg3 600 python bench.py --steps 100 --warmup 10 --sustain 3
run bench_cfg1 300 python bench.py --steps 200 --warmup 10 --workload cfg1 --no-cpu-baseline
run bench_cfg2 300 python bench.py --steps 200 --warmup 10 --workload cfg2 --no-cpu-baseline
run bench_cfg4 400 python bench.py --steps 50 --warmup 5 --workload cfg4 --no-cpu-baseline
run bench_big 400 python bench.py --steps 10 --warmup 3 --workload big --no-cpu-baseline
run bench_sweep 900 python bench.py --steps 50 --workload sweep
tail -n 5 gpurun_out/pytest.log gpurun_out/smoke.log
python tools/bench_summary.py gpurun_out/bench_*.log
cat gpurun_out/summary.txt
