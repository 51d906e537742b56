#!/bin/bash
# 8 x B200, final kernels: the headline scaling point and config 5
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 900 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/r2_bench_8gpu_final.json 2> gpurun_out/r2_bench_8gpu_final.err
cat gpurun_out/r2_bench_8gpu_final.json; tail -4 gpurun_out/r2_bench_8gpu_final.err
( time timeout 900 $TR --nproc-per-node 8 --master-port 29532 scripts/bench_config5.py ) > gpurun_out/r2_config5_8gpu_final.json 2> gpurun_out/r2_config5_8gpu_final.err
cat gpurun_out/r2_config5_8gpu_final.json; tail -4 gpurun_out/r2_config5_8gpu_final.err
