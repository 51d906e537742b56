#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_secondary.py > gpurun_out/secondary.jsonl 2> gpurun_out/secondary.err
cut -c1-260 gpurun_out/secondary.jsonl; tail -3 gpurun_out/secondary.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
