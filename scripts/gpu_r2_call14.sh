#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 600 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/r2_bench_8gpu_final2.json 2> gpurun_out/r2_bench_8gpu_final2.err
grep "^{" gpurun_out/r2_bench_8gpu_final2.json; tail -3 gpurun_out/r2_bench_8gpu_final2.err
timeout 300 $TR --nproc-per-node 8 --master-port 29552 scripts/check_sharded.py > gpurun_out/r2_check_sharded_8gpu.log 2>&1; tail -3 gpurun_out/r2_check_sharded_8gpu.log
