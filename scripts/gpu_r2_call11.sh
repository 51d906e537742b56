#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -k "knn or percentile or weights or fit or medium or config2 or pipeline_matches or duplicated" ) > gpurun_out/r2_pytest_gpu_e.log 2>&1
grep -E "passed|failed" gpurun_out/r2_pytest_gpu_e.log | tail -2
( time timeout 900 python scripts/bench_config5.py --genes-per-rank 3750 ) > gpurun_out/r2_config5_weak_1gpu.json 2> gpurun_out/r2_config5_weak_1gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_config5_weak_1gpu.json'))
print(d['stage_ms_max_over_ranks']); print(d['k5_csr']['frac_of_measured_hbm'])
PY
tail -3 gpurun_out/r2_config5_weak_1gpu.err
timeout 600 python scripts/bench_secondary.py 2>/dev/null | grep -E "knn_brute|percentile" | cut -c1-230
