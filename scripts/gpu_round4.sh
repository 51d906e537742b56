#!/bin/bash
mkdir -p gpurun_out
{
for args in "random 200 300" "random 1000 512" "random 4096 1024 100 700" "random 30000 384" "onehot" "time 30000 8192" "time 30000 16384"; do
  echo "=== v2 tc_selftest $args"; timeout 300 python scripts/tc_selftest.py $args 2>&1 | tail -18
done
echo "=== v1 time"; VELO_TC_VARIANT=1 timeout 300 python scripts/tc_selftest.py time 30000 8192 2>&1 | tail -3
} > gpurun_out/tc_selftest3.log 2>&1
cat gpurun_out/tc_selftest3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_coldeltacor_tc2 -c 1 -f -o gpurun_out/k2g_prof_v2 \
   python scripts/tc_selftest.py time 30000 4096 > gpurun_out/ncu_k2g_v2.log 2>&1
tail -3 gpurun_out/ncu_k2g_v2.log
