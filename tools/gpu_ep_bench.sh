#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_cfg3_ep${N}_s2.log 2>&1; echo "cfg3 exit=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --workload cfg4 > gpurun_out/bench_cfg4_ep${N}_s2.log 2>&1; echo "cfg4 exit=$?"
python tools/bench_summary.py gpurun_out/bench_cfg3_ep${N}_s2.log gpurun_out/bench_cfg4_ep${N}_s2.log | cut -c1-420
tail -3 gpurun_out/bench_cfg3_ep${N}_s2.log | grep -v "^{" | cut -c1-300
