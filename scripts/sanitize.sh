#!/usr/bin/env bash
# compute-sanitizer passes over small-shape GPU tests (memcheck on every kernel family, racecheck + synccheck on the
# shared-memory / mbarrier kernels).  Run on the GPU box:  bash scripts/sanitize.sh > gpurun_out/sanitize.log 2>&1
set -u
cd "$(dirname "$0")/.."
K1='golden_small or (ragged_and_multislab and 129) or (ragged_and_multislab and 130) or duplicate_indices or device_tier or fp32_ties'
K2='fit_slopes_match_reference_golden or knn_smoothing_matches_reference_golden or velocity_chain_matches or row_percentiles or sparse_counts or velocytoloom_pipeline'
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Error|error:|Invalid|Race|hazard" | tail -12; }
run compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_coldeltacor_gpu.py -q -x -k "$K1"
run compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_pipeline_gpu.py -q -x -k "$K2"
run compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_coldeltacor_gpu.py -q -x -k "golden_small or duplicate_indices or (ragged_and_multislab and 130)"
run compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_coldeltacor_gpu.py -q -x -k "golden_small or (ragged_and_multislab and 130)"
run compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_pipeline_gpu.py -q -x -k "row_percentiles or sparse_counts"
