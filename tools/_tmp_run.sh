timeout 900 python -m pytest tests/test_gpu_layer.py tests/test_gpu_ep.py -q -x --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -n 3 gpurun_out/pytest.log
timeout 200 python tools/route_trace.py 3200
for i in 1 2; do
  timeout 300 python bench.py --steps 200 --warmup 10 --workload cfg3 --no-cpu-baseline > gpurun_out/bench_cfg3.log 2>&1
  grep "^{" gpurun_out/bench_cfg3.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg3 us/layer', round(d['us_per_layer'],2), 'ffn', round(d['stage_us_per_layer']['expert_ffn'],2), 'route', round(d['stage_us_per_layer']['gate'],2), 'tok/s', f\"{d['value']:.3e}\")"
done
