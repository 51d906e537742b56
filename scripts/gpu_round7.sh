#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
( time timeout 1200 python bench.py ) > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err
cat gpurun_out/bench_main.json; tail -5 gpurun_out/bench_main.err
