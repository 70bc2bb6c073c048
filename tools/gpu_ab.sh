#!/bin/bash
# A/B of the launch-overlap / epilogue knobs on one box: parity first, then bench lines per variant.
#   variants: "PDL PDL_TRIG EPI2"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"
tail -n 3 gpurun_out/pytest.log
timeout 200 python tools/ffn_trace.py 3200 2 > gpurun_out/trace_3200.txt 2>&1; head -14 gpurun_out/trace_3200.txt
B200MOE_EPI2=0 timeout 200 python tools/ffn_trace.py 3200 2 > gpurun_out/trace_3200_epi0.txt 2>&1; head -14 gpurun_out/trace_3200_epi0.txt
timeout 200 python tools/ffn_trace.py 65536 1 > gpurun_out/trace_65536.txt 2>&1; head -14 gpurun_out/trace_65536.txt
rm -f gpurun_out/bench_*_v*.log
for v in "0 0 1" "0 0 0" "7 7 1" "7 3 1" "6 2 1" "4 0 1" "2 2 1"; do
  set -- $v
  for wl in ${WLS:-cfg3 cfg1 big}; do
    steps=100; [ $wl == big ] && steps=10
    B200MOE_PDL=$1 B200MOE_PDL_TRIG=$2 B200MOE_EPI2=$3 timeout 300 python bench.py --steps $steps --warmup 5 --workload $wl --no-cpu-baseline \
      > gpurun_out/bench_${wl}_v$1$2$3.log 2>&1
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*_v*.log')):
    ok=False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); ok=True
            print(f.split('/')[-1], 'us/layer', round(d['us_per_layer'],2), 'eager', round(d['ms_per_step_eager']*1e3/d['config']['layers'],2), 'stages', {k:(round(v,1) if v else v) for k,v in d['stage_us_per_layer'].items()}, 'tok/s', f"{d['value']:.3e}", 'e2e', f"{d['e2e']['value']:.3e}")
    if not ok: print(f, 'NO JSON'); print(open(f).read()[-1500:])
PY
