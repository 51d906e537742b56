#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29541 scripts/check_sharded.py > gpurun_out/r2_check_sharded_2gpu.log 2>&1; tail -5 gpurun_out/r2_check_sharded_2gpu.log
( time timeout 900 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu ) > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_2gpu.json'))
print(d['value'], d['clocks']); print(json.dumps(d['e2e']['variants']))
PY
tail -3 gpurun_out/r2_bench_2gpu.err
