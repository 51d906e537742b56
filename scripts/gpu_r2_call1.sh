#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -15 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python scripts/exact_variants.py > gpurun_out/r2_exact_variants.jsonl 2>&1
VELO_B200_LIB=$PWD/velocyto.py_b200/libvelo_b200_exact1024.so timeout 600 python scripts/exact_variants.py >> gpurun_out/r2_exact_variants.jsonl 2>&1
cat gpurun_out/r2_exact_variants.jsonl
( time timeout 1500 python bench.py ) > gpurun_out/r2_bench_main.json 2> gpurun_out/r2_bench_main.err
cat gpurun_out/r2_bench_main.json; tail -5 gpurun_out/r2_bench_main.err
