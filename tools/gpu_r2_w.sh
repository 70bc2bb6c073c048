#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for rev in 0 1 0 1; do
B200MOE_EP_P2REV=$rev timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_cfg3_ep${N}_rev$rev.log 2>&1; echo "rev $rev exit=$?"
python tools/bench_summary.py gpurun_out/bench_cfg3_ep${N}_rev$rev.log | cut -c1-200
done
