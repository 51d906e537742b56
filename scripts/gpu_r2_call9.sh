#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -k "sparse or knn_smoothing or config2 or medium" ) > gpurun_out/r2_pytest_gpu_c.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_c.log
timeout 300 python scripts/bench_config5.py --check > gpurun_out/r2_config5_check_1gpu.json 2> gpurun_out/r2_config5_check_1gpu.err; tail -c 300 gpurun_out/r2_config5_check_1gpu.json; tail -2 gpurun_out/r2_config5_check_1gpu.err
( time timeout 900 python scripts/bench_config5.py --genes-per-rank 3750 ) > gpurun_out/r2_config5_weak_1gpu.json 2> gpurun_out/r2_config5_weak_1gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_config5_weak_1gpu.json'))
print(d['stage_ms_max_over_ranks']); print(d['k5_csr']); print(d['properties'])
PY
tail -3 gpurun_out/r2_config5_weak_1gpu.err
