#!/bin/bash
# End-of-round validation on one B200: every -m gpu test, smoke(), the headline bench, one ncu traffic capture of K1 at the
# full configuration, the sanitizer passes.   gpurun --timeout 3000 -- 'bash scripts/gpu_final.sh'
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/r2_pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/r2_pytest_gpu.log | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1500 python bench.py ) > gpurun_out/r2_bench_main.json 2> gpurun_out/r2_bench_main.err
cat gpurun_out/r2_bench_main.json; tail -3 gpurun_out/r2_bench_main.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_coldeltacor -c 1 --clock-control none --csv --log-file gpurun_out/r2_k1_traffic_full_config.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-local > gpurun_out/r2_k1_traffic.log 2>&1
grep -E "dram__bytes|gpu__time" gpurun_out/r2_k1_traffic_full_config.csv | cut -c1-400
bash scripts/sanitize.sh > gpurun_out/r2_sanitize.log 2>&1; grep -E "===|passed|SUMMARY" gpurun_out/r2_sanitize.log | cut -c1-160
