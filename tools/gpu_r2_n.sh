#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_n.log 2>&1; echo "pytest exit=$?"
tail -n 4 gpurun_out/pytest_n.log
for pf in 2 0; do
  echo "=== prefetch $pf"
  B200MOE_FFN_PREFETCH=$pf python tools/timeline.py 3200 5 2>&1 | sed -n "/expert kernel, us/,\$p" | tail -n 4
  B200MOE_FFN_PREFETCH=$pf timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3_n_pf$pf.log 2>&1
  B200MOE_FFN_PREFETCH=$pf timeout 300 python bench.py --steps 200 --warmup 10 --workload cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1_n_pf$pf.log 2>&1
done
python tools/bench_summary.py gpurun_out/bench_cfg3_n_pf*.log gpurun_out/bench_cfg1_n_pf*.log | cut -c1-160
