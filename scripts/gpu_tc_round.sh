#!/bin/bash
# one gpurun call: K2g self-tests, timing, then the GPU test-suite
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
for args in "random 200 300" "onehot" "random 1000 512" "random 4096 1024 100 700" "random 30000 384" "time 30000 8192"; do
  echo "=== tc_selftest $args"; timeout 180 python scripts/tc_selftest.py $args 2>&1 | tail -40; echo "exit=$?"
done
} > gpurun_out/tc_selftest.log 2>&1
tail -60 gpurun_out/tc_selftest.log
if [ "$1" == "tests" ]; then
  ( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1
  tail -15 gpurun_out/pytest_gpu.log
fi
