#!/bin/bash
mkdir -p gpurun_out
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_coldeltacor -o gpurun_out/r2_k1_probe_final python scripts/k1_probe.py > gpurun_out/r2_k1_probe_final.log 2>&1
tail -2 gpurun_out/r2_k1_probe_final.log; ls -la gpurun_out/r2_k1_probe_final.ncu-rep
