#!/bin/bash
# static groups + weight prefetch before the dependency wait + fitted UMMA N: parity, then A/B per switch
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_g.log 2>&1; echo "pytest exit=$?"
tail -n 5 gpurun_out/pytest_g.log
for v in "1 1" "0 1" "1 0" "0 0"; do
  set -- $v
  B200MOE_FFN_PREFETCH=$1 B200MOE_FIT_N=$2 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3_pf$1_fn$2.log 2>&1
done
timeout 300 python bench.py --steps 200 --warmup 10 --workload cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1_g.log 2>&1
B200MOE_FFN_PREFETCH=0 B200MOE_FIT_N=0 timeout 300 python bench.py --steps 200 --warmup 10 --workload cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1_g00.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --workload big --no-cpu-baseline > gpurun_out/bench_big_g.log 2>&1
python tools/bench_summary.py gpurun_out/bench_cfg3_pf*.log gpurun_out/bench_cfg1_g*.log gpurun_out/bench_big_g.log
timeout 200 python tools/ffn_trace.py 3200 1 > gpurun_out/trace_3200_g.txt 2>&1; head -24 gpurun_out/trace_3200_g.txt
