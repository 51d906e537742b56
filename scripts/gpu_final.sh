#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_validate.sh
timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
   --clock-control none -k regex:^k_ --csv --log-file gpurun_out/secondary_ncu.csv python scripts/ncu_secondary.py > gpurun_out/secondary_ncu.log 2>&1
tail -2 gpurun_out/secondary_ncu.log; wc -l gpurun_out/secondary_ncu.csv
