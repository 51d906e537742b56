#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,smsp__cycles_active.avg --clock-control none -k regex:k_coldeltacor --csv --log-file gpurun_out/r2_k1_probe_v7.csv python scripts/k1_probe.py > gpurun_out/r2_k1_probe_v7.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2_k1_probe_v7.csv')) if len(r)>10]
h=rows[0]; by=collections.OrderedDict()
for r in rows[1:]: by.setdefault(r[h.index('ID')],{})[r[h.index('Metric Name')]]=r[h.index('Metric Value')]
for k,v in by.items(): print(k, v)
PY
timeout 600 python scripts/exact_variants.py > gpurun_out/r2_exact_variants_v7.jsonl 2>&1; cat gpurun_out/r2_exact_variants_v7.jsonl
( time timeout 1500 python bench.py ) > gpurun_out/r2_bench_main.json 2> gpurun_out/r2_bench_main.err
cat gpurun_out/r2_bench_main.json; tail -4 gpurun_out/r2_bench_main.err
