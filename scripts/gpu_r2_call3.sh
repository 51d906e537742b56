#!/bin/bash
# 2 x B200: fixed tests, API timings, sharded host-tier bench, config-5 check + weak-scaling points at N = 1, 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 900 python -m pytest tests -q -m gpu -k "medium or threshold or sparse_ingest or permute or host_tier or sharded_host" ) > gpurun_out/r2_pytest_gpu_b.log 2>&1
tail -8 gpurun_out/r2_pytest_gpu_b.log
timeout 600 python scripts/bench_pipeline.py > gpurun_out/r2_pipeline_config2_api_timings.json 2> gpurun_out/r2_pipeline_config2.err
cat gpurun_out/r2_pipeline_config2_api_timings.json | head -80; tail -3 gpurun_out/r2_pipeline_config2.err
timeout 300 $TR --nproc-per-node 2 --master-port 29511 scripts/check_sharded.py > gpurun_out/r2_check_sharded_2gpu.log 2>&1; tail -3 gpurun_out/r2_check_sharded_2gpu.log
timeout 300 $TR --nproc-per-node 2 --master-port 29512 scripts/bench_config5.py --check > gpurun_out/r2_config5_check_2gpu.json 2> gpurun_out/r2_config5_check_2gpu.err; tail -c 600 gpurun_out/r2_config5_check_2gpu.json; tail -3 gpurun_out/r2_config5_check_2gpu.err
( time timeout 900 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
cat gpurun_out/r2_bench_2gpu.json; tail -5 gpurun_out/r2_bench_2gpu.err
( time timeout 900 python scripts/bench_config5.py --genes-per-rank 3750 ) > gpurun_out/r2_config5_weak_1gpu.json 2> gpurun_out/r2_config5_weak_1gpu.err
cat gpurun_out/r2_config5_weak_1gpu.json; tail -5 gpurun_out/r2_config5_weak_1gpu.err
( time timeout 900 $TR --nproc-per-node 2 --master-port 29514 scripts/bench_config5.py --genes-per-rank 3750 ) > gpurun_out/r2_config5_weak_2gpu.json 2> gpurun_out/r2_config5_weak_2gpu.err
cat gpurun_out/r2_config5_weak_2gpu.json; tail -5 gpurun_out/r2_config5_weak_2gpu.err
