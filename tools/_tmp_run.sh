timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -n 4 gpurun_out/pytest.log
for kp in 2 1; do
for wl in big cfg3 cfg1; do
  steps=100; [ $wl == big ] && steps=10
  B200MOE_KPS=$kp timeout 300 python bench.py --steps $steps --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/bench_${wl}_k$kp.log 2>&1
  grep "^{" gpurun_out/bench_${wl}_k$kp.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl kps=$kp us/layer', round(d['us_per_layer'],2), 'stages', {k:(round(v,1) if v else v) for k,v in d['stage_us_per_layer'].items()}, 'roof', d['roofline']['bound'], round(d['roofline']['frac'],3), 'tok/s', f\"{d['value']:.3e}\")" || tail -5 gpurun_out/bench_${wl}_k$kp.log
done; done
timeout 200 python tools/ffn_trace.py 65536 1 > gpurun_out/trace_65536.txt 2>&1; head -12 gpurun_out/trace_65536.txt
