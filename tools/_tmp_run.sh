timeout 600 python -m pytest tests/test_gpu_ep.py -q -x --timeout 200 > gpurun_out/pytest_ep.log 2>&1; echo "pytest_ep exit=$?"; tail -n 3 gpurun_out/pytest_ep.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ep_check.py > gpurun_out/ep_check.log 2>&1; echo "ep_check exit=$?"
for i in 1 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench2 exit=$?"; grep "^{" gpurun_out/bench_n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2 us/layer', d['us_per_layer'], 'eager', d['ms_per_step_eager']*1e3/18, 'tok/s', d['value'], 'e2e', d['e2e']['value'])"
done
