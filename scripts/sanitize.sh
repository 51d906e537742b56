#!/usr/bin/env bash
# compute-sanitizer passes over small-shape GPU tests: memcheck on every kernel family, racecheck + synccheck on the
# shared-memory / mbarrier / cluster kernels.  Run on the GPU box:  bash scripts/sanitize.sh > gpurun_out/sanitize.log 2>&1
# Round 2 adds the families round 1 never ran under the tools: the tcgen05 CTA-pair kernel (cluster mbarriers, TMEM),
# random.cu (TMA-staged row permutation, per-cell sort), preprocess.cu, the staged / chunk-pipelined host tier,
# the kNN kernel incl. its tie path, sparse ingest, grid flow.
set -u
cd "$(dirname "$0")/.."
K1='golden_small or (ragged_and_multislab and 129) or (ragged_and_multislab and 130) or duplicate_indices or device_tier or fp32_ties or small_sigma'
K2='fit_slopes_match_reference_golden or knn_smoothing_matches_reference_golden or velocity_chain_matches or row_percentiles or sparse_counts or velocytoloom_pipeline_matches'
TC='(tensor_core_linear_matches_oracle and 777) or (tensor_core_linear_matches_oracle and 64-128) or tensor_core_linear_degenerate'
HOST='host_tier_pipelined or host_tier_pageable or sharded_host_entry'
RND='device_permute_rows_nsign or neighbour_sampler or normalize_matches_reference_golden or duplicated_points or (angular and correlation) or sparse_ingest or velocity_threshold'
run() { echo "=== $*"; timeout 1500 "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|error:|Invalid|Race|hazard" | tail -12; }
run compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_coldeltacor_gpu.py -q -x -k "$K1"
run compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_pipeline_gpu.py -q -x -k "$K2"
run compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_coldeltacor_gpu.py -q -x -k "$TC"
run compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_coldeltacor_gpu.py -q -x -k "$HOST"
run compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_pipeline_gpu.py -q -x -k "$RND"
run compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_coldeltacor_gpu.py -q -x -k "golden_small or duplicate_indices or (ragged_and_multislab and 130)"
run compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_coldeltacor_gpu.py -q -x -k "$TC"
run compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_pipeline_gpu.py -q -x -k "row_percentiles or sparse_counts or device_permute_rows_nsign or neighbour_sampler or duplicated_points"
run compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_coldeltacor_gpu.py -q -x -k "golden_small or (ragged_and_multislab and 130)"
run compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_coldeltacor_gpu.py -q -x -k "$TC"
run compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_pipeline_gpu.py -q -x -k "row_percentiles or device_permute_rows_nsign or neighbour_sampler or duplicated_points"
