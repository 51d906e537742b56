#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python scripts/bench_config3.py > gpurun_out/bench_config3.jsonl 2> gpurun_out/bench_config3.err
cat gpurun_out/bench_config3.jsonl; tail -3 gpurun_out/bench_config3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_k2g.csv \
   python scripts/tc_selftest.py time 30000 4096 > gpurun_out/launches_k2g.log 2>&1
tail -2 gpurun_out/launches_k2g.log
