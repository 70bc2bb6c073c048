#!/bin/bash
# ncu launch list and per-kernel DRAM / tensor metrics of this repository's kernels only (regex on the namespace)
mkdir -p gpurun_out
O=gpurun_out
CFG3="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum
$CFG3 > $O/final_ncu_plain.log 2>&1 && {
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'route_kernel|ffn_kernel|gate_|dispatch_|combine_|layernorm' -s 80 -c 108 --csv --log-file $O/final_launches_cfg3.csv $CFG3 > $O/final_ncu_list.log 2>&1
  echo "ncu_list exit=$?"
  ncu --metrics $M --clock-control none -k regex:'route_kernel|ffn_kernel|gate_|dispatch_|combine_|layernorm' -s 80 -c 72 --csv --log-file $O/final_metrics_cfg3.csv $CFG3 > $O/final_ncu_metrics.log 2>&1
  echo "ncu_metrics exit=$?"
}
BLK="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --block"
$BLK > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'route_kernel|ffn_kernel|gate_|dispatch_|combine_|layernorm' -s 80 -c 108 --csv --log-file $O/final_launches_cfg3_block.csv $BLK > $O/final_ncu_list_block.log 2>&1
echo "ncu_list_block exit=$?"
python tools/ncu_summary.py launches $O/final_launches_cfg3.csv
python tools/ncu_summary.py launches $O/final_launches_cfg3_block.csv
