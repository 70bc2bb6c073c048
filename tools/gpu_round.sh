#!/bin/bash
# Routine GPU pass: parity suite, smoke, bench, then an ncu launch list and one full capture of the FFN kernel.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1; local to=$2; shift 2; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; }
run pytest 900 python -m pytest tests -m gpu -q --timeout 300
run smoke  200 python __graft_entry__.py smoke
run bench  600 python bench.py --steps 100 --warmup 10
run bench_cfg1 300 python bench.py --steps 200 --warmup 10 --workload cfg1 --no-cpu-baseline
run bench_big 400 python bench.py --steps 10 --warmup 3 --workload big --no-cpu-baseline
if [ "$1" == "ncu" ]; then
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_list.log 2>&1
  echo "ncu_list exit=$?" | tee -a gpurun_out/summary.txt
  ncu --set full --clock-control none --import-source on -k regex:ffn_kernel -s 60 -c 3 -o gpurun_out/ffn_prof -f \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
  echo "ncu_full exit=$?" | tee -a gpurun_out/summary.txt
fi
tail -n 4 gpurun_out/pytest.log gpurun_out/smoke.log
grep -h '^{' gpurun_out/bench*.log
cat gpurun_out/summary.txt
