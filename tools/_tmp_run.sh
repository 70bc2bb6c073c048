timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -n 5 gpurun_out/pytest.log
timeout 200 python tools/route_trace.py 3200
timeout 200 python tools/route_trace.py 50
for r in 1; do
for wl in cfg3 cfg1; do
  B200MOE_ROUTE=$r timeout 300 python bench.py --steps 100 --warmup 5 --workload $wl --no-cpu-baseline > gpurun_out/bench_${wl}_r$r.log 2>&1
  grep "^{" gpurun_out/bench_${wl}_r$r.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl route=$r us/layer', round(d['us_per_layer'],2), 'eager', round(d['ms_per_step_eager']*1e3/d['config']['layers'],2), 'stages', {k:(round(v,1) if v else v) for k,v in d['stage_us_per_layer'].items()}, 'tok/s', f\"{d['value']:.3e}\", 'e2e', f\"{d['e2e']['value']:.3e}\")" || tail -5 gpurun_out/bench_${wl}_r$r.log
done; done
