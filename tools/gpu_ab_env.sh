#!/bin/bash
# A/B of environment knobs of ONE library build on one box (boxes differ by +-5 %): parity of the variants first, then
# bench lines alternating between the variants, twice.
#   (a variant may set several variables, comma separated: A=1,B=2)
#   usage: VARIANTS="B200MOE_2PROD=0 B200MOE_2PROD=1" WLS="cfg3 cfg2" tools/gpu_ab_env.sh [pytest -k expression]
mkdir -p gpurun_out
VARIANTS=${VARIANTS:-"B200MOE_2PROD=0 B200MOE_2PROD=1"}
WLS=${WLS:-cfg3 cfg2 cfg4}
K=${1:-"expert_ffn or layer or ep"}
i=0
for v in $VARIANTS; do
  if [ $i -gt 0 ]; then   # (the first variant is the shipped default: covered by the full suite elsewhere)
    env ${v//,/ } timeout 600 python -m pytest tests -m gpu -q -x --timeout 200 -k "$K" > gpurun_out/abenv_pytest_$i.log 2>&1
    echo "$v pytest exit=$?"; tail -n 2 gpurun_out/abenv_pytest_$i.log
  fi
  i=$((i + 1))
done
rm -f gpurun_out/abenv_*_v*.log
for round in 1 2; do
  i=0
  for v in $VARIANTS; do
    for wl in $WLS; do
      steps=200; [ $wl == big ] && steps=10
      env ${v//,/ } timeout 300 python bench.py --steps $steps --warmup 10 --workload $wl --no-cpu-baseline > gpurun_out/abenv_${wl}_v${i}_$round.log 2>&1
    done
    i=$((i + 1))
  done
done
for wl in $WLS; do python tools/bench_summary.py gpurun_out/abenv_${wl}_v*.log | cut -c1-200; done
