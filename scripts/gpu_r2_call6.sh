#!/bin/bash
# 8 x B200: the scaling points that need the whole box
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 900 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
cat gpurun_out/r2_bench_8gpu.json; tail -4 gpurun_out/r2_bench_8gpu.err
( time timeout 900 $TR --nproc-per-node 8 --master-port 29522 scripts/bench_config5.py ) > gpurun_out/r2_config5_8gpu.json 2> gpurun_out/r2_config5_8gpu.err
cat gpurun_out/r2_config5_8gpu.json; tail -4 gpurun_out/r2_config5_8gpu.err
timeout 300 $TR --nproc-per-node 8 --master-port 29523 scripts/bench_config5.py --check > gpurun_out/r2_config5_check_8gpu.json 2> gpurun_out/r2_config5_check_8gpu.err; tail -c 500 gpurun_out/r2_config5_check_8gpu.json; tail -2 gpurun_out/r2_config5_check_8gpu.err
timeout 300 $TR --nproc-per-node 8 --master-port 29524 scripts/check_sharded.py > gpurun_out/r2_check_sharded_8gpu.log 2>&1; tail -3 gpurun_out/r2_check_sharded_8gpu.log
( time timeout 600 $TR --nproc-per-node 4 --master-port 29525 scripts/bench_config5.py --genes-per-rank 3750 ) > gpurun_out/r2_config5_weak_4gpu.json 2> gpurun_out/r2_config5_weak_4gpu.err
cut -c1-1200 gpurun_out/r2_config5_weak_4gpu.json; tail -2 gpurun_out/r2_config5_weak_4gpu.err
( time timeout 600 $TR --nproc-per-node 2 --master-port 29526 scripts/bench_config5.py --genes-per-rank 3750 ) > gpurun_out/r2_config5_weak_2gpu.json 2> gpurun_out/r2_config5_weak_2gpu.err
cut -c1-1200 gpurun_out/r2_config5_weak_2gpu.json; tail -2 gpurun_out/r2_config5_weak_2gpu.err
( time timeout 600 $TR --nproc-per-node 4 --master-port 29527 bench.py --gpus 4 --steps 3 --warmup 3 ) > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err
cut -c1-2500 gpurun_out/r2_bench_4gpu.json; tail -2 gpurun_out/r2_bench_4gpu.err
