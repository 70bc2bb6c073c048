#!/bin/bash
# Routine GPU pass: parity suite, smoke, bench lines; optional ncu launch list + full capture of one kernel.
#   bash tools/gpu_round.sh [ncu [kernel-regex]]
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1; local to=$2; shift 2; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; }
run pytest 900 python -m pytest tests -m gpu -q --timeout 300
run smoke  200 python __graft_entry__.py smoke
run bench  600 python bench.py --steps 100 --warmup 10
run bench_cfg1 300 python bench.py --steps 200 --warmup 10 --workload cfg1 --no-cpu-baseline
run bench_big 400 python bench.py --steps 10 --warmup 3 --workload big --no-cpu-baseline
if [ "$1" == "ncu" ]; then
  KREG=${2:-ffn_kernel}
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_list.log 2>&1
  echo "ncu_list exit=$?" | tee -a gpurun_out/summary.txt
  ncu --set full --clock-control none --import-source on -k regex:$KREG -s 60 -c 2 -o gpurun_out/prof_$KREG -f \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
  echo "ncu_full exit=$?" | tee -a gpurun_out/summary.txt
fi
tail -n 4 gpurun_out/pytest.log gpurun_out/smoke.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench*.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f, d['config']['workload'][:5], 'us/layer', round(d['us_per_layer'],2), 'stages', {k:(round(v,1) if v else v) for k,v in d['stage_us_per_layer'].items()}, 'roof', d['roofline']['bound'], round(d['roofline']['frac'],3), 'tok/s', f"{d['value']:.3e}", 'e2e', f"{d['e2e']['value']:.3e}", 'cpu', d.get('cpu_baseline') and f"{d['cpu_baseline']['value']:.3e}")
PY
cat gpurun_out/summary.txt
