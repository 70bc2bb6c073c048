#!/bin/bash
# First-contact script for a fresh B200 box: every group runs in its own process under a timeout so that one hung or
# faulting kernel cannot take the rest of the run with it. Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, command...
  local name=$1; local to=$2; shift 2
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  echo "$name exit=$?" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run gate      300 python -m pytest tests/test_gpu_stages.py -q -k "gate or softmax" --timeout 120
run dispatch  300 python -m pytest tests/test_gpu_stages.py -q -k "dispatch" --timeout 120
run combine   200 python -m pytest tests/test_gpu_stages.py -q -k "combine" --timeout 120
run ffn       400 python -m pytest tests/test_gpu_stages.py -q -k "expert_ffn" --timeout 120
run layer     600 python -m pytest tests/test_gpu_layer.py -q --timeout 200
run smoke     200 python __graft_entry__.py smoke
run bench     600 python bench.py --steps 50 --warmup 5
tail -n 5 gpurun_out/*.log
cat gpurun_out/summary.txt
