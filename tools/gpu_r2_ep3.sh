#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests/test_gpu_ep.py -q -x --timeout 120 2>&1 | tail -3
timeout 600 $TR tools/ep_check.py > gpurun_out/ep_check_n$N.log 2>&1; echo "ep_check exit=$?"; grep -c " ok" gpurun_out/ep_check_n$N.log; grep -E "FAIL|Error|error" gpurun_out/ep_check_n$N.log | head
timeout 300 $TR tools/ep_trace.py 3200 > gpurun_out/ep_trace_n$N.txt 2>&1; echo "ep_trace exit=$?"; grep -v "^\*\|NCCL\|W1018\|^$" gpurun_out/ep_trace_n$N.txt | head -34
timeout 600 $TR bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_cfg3_ep$N.log 2>&1; echo "bench cfg3 exit=$?"
timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 --workload cfg4 > gpurun_out/bench_cfg4_ep$N.log 2>&1; echo "bench cfg4 exit=$?"
python tools/bench_summary.py gpurun_out/bench_cfg3_ep$N.log gpurun_out/bench_cfg4_ep$N.log
