#!/bin/bash
# Quick perf iteration: targeted parity tests, FFN timeline, bench lines.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_layer.py -q -x --timeout 200 -k "${1:-ffn or layer or golden}" > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit=$?"
tail -n 3 gpurun_out/pytest_quick.log
timeout 200 python tools/ffn_trace.py 3200 2 > gpurun_out/trace_3200.txt 2>&1; python tools/trace_stats.py gpurun_out/ffn_trace_3200.npy
timeout 200 python tools/ffn_trace.py 50 1 > gpurun_out/trace_50.txt 2>&1; python tools/trace_stats.py gpurun_out/ffn_trace_50.npy
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --steps 200 --warmup 10 --workload cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --workload big --no-cpu-baseline > gpurun_out/bench_big.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench*.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f, d['config']['workload'][:5], 'us/layer', round(d['us_per_layer'],2), 'stages', {k:(round(v,1) if v else v) for k,v in d['stage_us_per_layer'].items()}, 'roof', d['roofline']['bound'], round(d['roofline']['frac'],3), 'tok/s', f"{d['value']:.3e}", 'e2e', f"{d['e2e']['value']:.3e}")
PY
