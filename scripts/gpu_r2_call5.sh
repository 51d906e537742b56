#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -12 gpurun_out/r2_pytest_gpu.log
PROFILE=1 SKIP_REFERENCE_RNG=1 timeout 600 python scripts/bench_pipeline.py > gpurun_out/r2_pipeline_config2_profiled.json 2> gpurun_out/r2_pipeline_config2_profiled.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_pipeline_config2_profiled.json'))
for k,v in d['methods'].items(): print(f"{v['wall_ms']:10.1f} ms  {k}")
PY
grep -A 30 "cumulative" gpurun_out/r2_pipeline_config2_profiled.err | cut -c1-160 | head -40
timeout 900 python scripts/bench_config3.py --skip-fp32 > gpurun_out/r2_bench_config3.jsonl 2> gpurun_out/r2_bench_config3.err
cut -c1-700 gpurun_out/r2_bench_config3.jsonl; tail -3 gpurun_out/r2_bench_config3.err
timeout 600 python scripts/bench_config3.py --cells 8192 --skip-fp32 > gpurun_out/r2_bench_config3_8k.jsonl 2>> gpurun_out/r2_bench_config3.err
cut -c1-400 gpurun_out/r2_bench_config3_8k.jsonl
