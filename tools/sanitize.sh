#!/bin/bash
# compute-sanitizer over the small fdPA workload with the concurrency lanes on (race / memory checks).
# memcheck: out-of-bounds and misaligned accesses, leaks of the context; racecheck: shared-memory hazards inside the kernels
# (slab_conv / slab_own / q-lane entry staging); initcheck: reads of device memory no kernel or copy has written
# (e.g. a lane reading a scratch table before its producer ran).  Logs -> gpurun_out/sanitize_*.log
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|checksum" gpurun_out/sanitize_$tool.log | tail -4
done
