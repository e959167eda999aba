#!/bin/bash
# ncu --set full of the s-wave solver's kernels over two steps of `bench.py --nl-method 1` (eager issue), raw page as CSV
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:"sw_|swave_tables_nl" -c 40 -o /tmp/swave python bench.py --nl-method 1 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/swave_ncu.log 2>&1
ncu -i /tmp/swave.ncu-rep --page raw --csv > gpurun_out/swave_raw.csv 2>/dev/null
ls -la gpurun_out/swave_raw.csv
