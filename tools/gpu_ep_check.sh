#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_ep.py -q -x --timeout 200 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/ep_check.py > gpurun_out/ep_check_n$N.log 2>&1; echo "ep_check exit=$?"; grep -c " ok" gpurun_out/ep_check_n$N.log; grep -E "FAIL|Error|error|Traceback" gpurun_out/ep_check_n$N.log | head -5
timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_cfg3_ep${N}_seg.log 2>&1; echo "cfg3 exit=$?"
timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 --workload cfg4 > gpurun_out/bench_cfg4_ep${N}_seg.log 2>&1; echo "cfg4 exit=$?"
python tools/bench_summary.py gpurun_out/bench_cfg3_ep${N}_seg.log gpurun_out/bench_cfg4_ep${N}_seg.log | cut -c1-420
timeout 300 $TR tools/timeline_ep.py 3200 4 2>&1 | grep -v "^\*\|NCCL\|OMP\|^$" | tail -12
