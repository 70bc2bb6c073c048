#!/bin/bash
# Round-2 evidence on ONE GPU: parity suite, smoke, every bench workload, reference arm, ncu launch list / full capture /
# per-kernel DRAM metrics.  Outputs in gpurun_out/final_*.  SHORT=1 leaves out the sweep, the route kernel's full capture
# and the micro-benchmarks (when only the expert kernel changed).
mkdir -p gpurun_out
O=gpurun_out
: > $O/final_summary.txt
run() { local name=$1; local to=$2; shift 2; timeout "$to" "$@" > "$O/final_$name.log" 2>&1; echo "$name exit=$?" | tee -a $O/final_summary.txt; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > $O/final_gpu.txt 2>&1
run pytest 1200 python -m pytest tests -m gpu -q --timeout 300
run smoke  200 python __graft_entry__.py smoke
run bench_cfg3 900 python bench.py --steps 200 --warmup 10 --sustain 3 --encoder
run bench_ref 600 python bench.py --impl reference --steps 3 --warmup 1
run bench_cfg1 300 python bench.py --steps 300 --warmup 10 --workload cfg1 --no-cpu-baseline
run bench_cfg2 400 python bench.py --steps 300 --warmup 10 --workload cfg2 --no-cpu-baseline --encoder
run bench_cfg3_block 300 python bench.py --steps 200 --warmup 10 --block --no-cpu-baseline
run bench_cfg3f 300 python bench.py --steps 100 --warmup 10 --workload cfg3f --no-cpu-baseline
run bench_cfg3_tf32 300 python bench.py --steps 100 --warmup 10 --compute tf32 --no-cpu-baseline
run bench_big 400 python bench.py --steps 10 --warmup 3 --workload big --no-cpu-baseline
run bench_cfg4 400 python bench.py --steps 20 --warmup 3 --workload cfg4 --no-cpu-baseline
[ -z "$SHORT" ] && run bench_sweep 900 python bench.py --workload sweep --steps 20 --warmup 3
CFG3="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum
$CFG3 > $O/final_ncu_plain.log 2>&1 && {
  ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 100 --csv --log-file $O/final_launches_cfg3.csv $CFG3 > $O/final_ncu_list.log 2>&1
  echo "ncu_list exit=$?" | tee -a $O/final_summary.txt
  ncu --metrics $M --clock-control none -s 120 -c 72 --csv --log-file $O/final_metrics_cfg3.csv $CFG3 > $O/final_ncu_metrics.log 2>&1
  echo "ncu_metrics exit=$?" | tee -a $O/final_summary.txt
  ncu --set full --clock-control none --import-source on -k regex:ffn_kernel -s 40 -c 2 -f -o $O/final_ffn_cfg3 $CFG3 > $O/final_ncu_full.log 2>&1
  echo "ncu_full exit=$?" | tee -a $O/final_summary.txt
  [ -z "$SHORT" ] && ncu --set full --clock-control none --import-source on -k regex:route_kernel -s 40 -c 2 -f -o $O/final_route_cfg3 $CFG3 > $O/final_ncu_full_route.log 2>&1
  echo "ncu_full_route exit=$?" | tee -a $O/final_summary.txt
}
python tools/timeline.py 3200 6 > $O/final_timeline_cfg3.txt 2>&1
python tools/route_trace.py 3200 > $O/final_route_trace.txt 2>&1
[ -z "$SHORT" ] && tools/bin/mem_facts_bench > $O/final_mem_facts.txt 2>&1
[ -z "$SHORT" ] && tools/bin/mma_rate_bench > $O/final_mma_rate.txt 2>&1
tail -n 3 $O/final_pytest.log $O/final_smoke.log
python tools/bench_summary.py $O/final_bench_*.log | cut -c1-360
cat $O/final_summary.txt
