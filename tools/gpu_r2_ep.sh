#!/bin/bash
# multi-GPU pass: parity across real GPUs, then the bench lines.  usage: bash tools/gpu_r2_ep.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/ep_check.py > gpurun_out/ep_check_n$N.log 2>&1; echo "ep_check exit=$?"; grep -E "ok|FAIL|Error|error" gpurun_out/ep_check_n$N.log | tail -40
timeout 600 $TR bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_cfg3_ep$N.log 2>&1; echo "bench cfg3 exit=$?"
timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 --workload cfg4 > gpurun_out/bench_cfg4_ep$N.log 2>&1; echo "bench cfg4 exit=$?"
B200MOE_EP_FOLD=0 timeout 600 $TR bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_cfg3_ep${N}_nofold.log 2>&1; echo "bench cfg3 nofold exit=$?"
python tools/bench_summary.py gpurun_out/bench_cfg3_ep$N.log gpurun_out/bench_cfg4_ep$N.log gpurun_out/bench_cfg3_ep${N}_nofold.log
tail -3 gpurun_out/bench_cfg3_ep$N.log | cut -c1-600
