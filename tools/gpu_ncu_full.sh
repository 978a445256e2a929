#!/bin/bash
# Per-kernel DRAM traffic / pipe shares of one eager step with a SHORT metric list (a --set full pass over the ~140 launches took
# 11 minutes and a 149 MB report); only the exported raw page (CSV) is kept.
TAG=${1:-r02}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off -o /tmp/${TAG}_step python tools/ncu_step.py > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ncu -i /tmp/${TAG}_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_step_raw.csv /tmp/${TAG}_step.ncu-rep
