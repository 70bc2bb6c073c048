#!/bin/bash
# Round-2 ncu evidence (one GPU): compute-bound FFN (full set), per-kernel DRAM bytes / tensor activity of every kernel
# of the path at cfg3 and at a 65 536-token top-2 sweep point (gate_tc, dispatch, combine).
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
BIG="python bench.py --workload big --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
CFG3="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
K2="python bench.py --workload sweep --sweep-points 65536:2 --steps 3 --no-graph"
$BIG > gpurun_out/ncu_plain_big.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ffn_kernel -s 40 -c 2 -f -o gpurun_out/r02_ffn_big $BIG > gpurun_out/ncu_big.log 2>&1
echo "ncu big exit=$?"
$CFG3 > gpurun_out/ncu_plain_cfg3.log 2>&1 && \
ncu --metrics $M --clock-control none -s 120 -c 72 --csv --log-file gpurun_out/r02_metrics_cfg3.csv $CFG3 > gpurun_out/ncu_cfg3.log 2>&1
echo "ncu cfg3 exit=$?"
$K2 > gpurun_out/ncu_plain_k2.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:'gate_tc|dispatch|combine|ffn_kernel' -s 40 -c 32 --csv --log-file gpurun_out/r02_metrics_65536_k2.csv $K2 > gpurun_out/ncu_k2.log 2>&1
echo "ncu k2 exit=$?"
$BIG > /dev/null 2>&1 && \
ncu --metrics $M --clock-control none -k regex:'gate_tc|dispatch|ffn_kernel' -s 40 -c 24 --csv --log-file gpurun_out/r02_metrics_big.csv $BIG > gpurun_out/ncu_big2.log 2>&1
echo "ncu big metrics exit=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_metrics*.csv
