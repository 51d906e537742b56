#!/bin/bash
mkdir -p gpurun_out
{
for args in "random 200 300" "random 4096 1024 100 700" "bias 30000 256" "time 30000 8192" "time 30000 16384"; do
  echo "=== BK32 tc_selftest $args"; timeout 300 python scripts/tc_selftest.py $args 2>&1 | tail -12
done
echo "=== BK64 time"; VELO_B200_LIB=$PWD/velocyto.py_b200/libvelo_b200_bk64.so timeout 300 python scripts/tc_selftest.py time 30000 8192 2>&1 | tail -3
} > gpurun_out/tc_selftest2.log 2>&1
cat gpurun_out/tc_selftest2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_coldeltacor_tc -c 1 -f -o gpurun_out/k2g_prof_bk32 \
   python scripts/tc_selftest.py time 30000 4096 > gpurun_out/ncu_k2g_bk32.log 2>&1
tail -3 gpurun_out/ncu_k2g_bk32.log
