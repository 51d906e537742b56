#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -12 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python scripts/bench_pipeline.py > gpurun_out/r2_pipeline_config2_api_timings.json 2> gpurun_out/r2_pipeline_config2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_pipeline_config2_api_timings.json'))
for k,v in d['methods'].items(): print(f"{v['wall_ms']:10.1f} ms  {k}")
PY
C=100000 G=5000 K=500 NN=10000 timeout 900 python scripts/bench_pipeline.py > gpurun_out/r2_pipeline_100k_api_timings.json 2> gpurun_out/r2_pipeline_100k.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_pipeline_100k_api_timings.json'))
for k,v in d['methods'].items(): print(f"{v['wall_ms']:10.1f} ms  {k}")
PY
tail -3 gpurun_out/r2_pipeline_100k.err
( time timeout 900 python scripts/bench_config5.py --genes-per-rank 3750 ) > gpurun_out/r2_config5_weak_1gpu.json 2> gpurun_out/r2_config5_weak_1gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_config5_weak_1gpu.json'))
print(d['stage_ms_max_over_ranks']); print(d['k5_csr']); print(d['memory'])
PY
tail -3 gpurun_out/r2_config5_weak_1gpu.err
timeout 900 python scripts/bench_secondary.py 2>/dev/null | grep -E "permute|percentile|knn_brute|K5" | cut -c1-200 > gpurun_out/r2_secondary_b.jsonl; cat gpurun_out/r2_secondary_b.jsonl
( time timeout 1500 python bench.py --no-local ) > gpurun_out/r2_bench_main.json 2> gpurun_out/r2_bench_main.err
cat gpurun_out/r2_bench_main.json; tail -4 gpurun_out/r2_bench_main.err
bash scripts/sanitize.sh > gpurun_out/r2_sanitize.log 2>&1; cat gpurun_out/r2_sanitize.log
