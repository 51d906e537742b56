#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded.py > gpurun_out/check_sharded_$N.log 2>&1
tail -6 gpurun_out/check_sharded_$N.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 ) > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
cat gpurun_out/bench_${N}gpu.json; tail -4 gpurun_out/bench_${N}gpu.err
