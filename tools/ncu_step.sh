#!/bin/bash
# ncu --set full of the main kernels of one bench step; exports the raw and source pages as CSV and drops the (large) report.
# usage: tools/ncu_step.sh <tag> <kernel regex> <count>
tag=$1; rx=$2; cnt=${3:-20}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -o /tmp/$tag python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/$tag.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_source_sass.csv 2>/dev/null
ls -la /tmp/$tag.ncu-rep gpurun_out/${tag}_*
