#!/bin/bash
# round 2, GPU session A: parity tests, A/B of the q-lane kernel against the column kernel, ncu of the new kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
FDGA_QLANE=0 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_column.json 2> gpurun_out/r2a_bench_column.err
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_qlane.json 2> gpurun_out/r2a_bench_qlane.err
python - <<'PY'
import json
for n in ("column","qlane"):
    try:
        d=json.loads(open(f"gpurun_out/r2a_bench_{n}.json").read().strip().splitlines()[-1]); k=d["kernels"]
        print(n, round(d["value"],1), round(d["e2e"]["value"],1), d["state_sha1"], {x:round(k[x]["ms_per_step"],3) for x in k})
    except Exception as e: print(n, "failed", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2a_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qlane_kernel -c 12 -o gpurun_out/r2a_qlane python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2a_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
