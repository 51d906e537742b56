#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -40 gpurun_out/r2_pytest_gpu.log
timeout 600 python scripts/bench_config5.py --check > gpurun_out/r2_config5_check_1gpu.json 2> gpurun_out/r2_config5_check_1gpu.err
tail -c 1500 gpurun_out/r2_config5_check_1gpu.json; tail -5 gpurun_out/r2_config5_check_1gpu.err
timeout 900 python scripts/bench_secondary.py > gpurun_out/r2_secondary.jsonl 2> gpurun_out/r2_secondary.err
cut -c1-220 gpurun_out/r2_secondary.jsonl; tail -3 gpurun_out/r2_secondary.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2_secondary_ncu.csv python scripts/ncu_secondary.py > gpurun_out/r2_secondary_ncu.log 2>&1
tail -2 gpurun_out/r2_secondary_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_coldeltacor -o gpurun_out/r2_k1_probe python scripts/k1_probe.py > gpurun_out/r2_k1_probe.log 2>&1
tail -3 gpurun_out/r2_k1_probe.log; ls -la gpurun_out/*.ncu-rep
