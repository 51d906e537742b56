#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_coldeltacor_tc -c 1 -f -o gpurun_out/k2g_prof \
   python scripts/tc_selftest.py time 30000 4096 > gpurun_out/ncu_k2g.log 2>&1
tail -5 gpurun_out/ncu_k2g.log
ls -la gpurun_out
