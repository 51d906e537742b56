#!/bin/bash
# One gpurun call that re-validates the round on one B200: GPU parity tests, smoke(), and the headline bench.
#   gpurun --timeout 1800 -- 'bash scripts/gpu_validate.sh'
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1200 python bench.py "$@" ) > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err
cat gpurun_out/bench_main.json; tail -4 gpurun_out/bench_main.err
