#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python scripts/bench_config3.py > gpurun_out/bench_config3.jsonl 2> gpurun_out/bench_config3.err
cut -c1-400 gpurun_out/bench_config3.jsonl; tail -3 gpurun_out/bench_config3.err
timeout 300 python scripts/tc_selftest.py time 30000 4096 2>&1 | tail -3
( time timeout 1200 python bench.py ) > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err
cat gpurun_out/bench_main.json; tail -5 gpurun_out/bench_main.err
