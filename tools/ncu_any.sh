#!/bin/bash
# tools/ncu_any.sh <tag> <kernel regex> <count> <bench args...>: ncu --set full of selected kernels, raw page exported as CSV
tag=$1; rx=$2; cnt=$3; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:"$rx" -c $cnt -o /tmp/$tag python bench.py "$@" --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_raw.csv
