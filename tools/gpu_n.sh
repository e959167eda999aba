#!/bin/bash
# multi-GPU session: tools/gpu_n.sh <N> <tag> "<extra bench args>" ...   (one bench run per extra-args string)
N=$1; tag=$2; shift 2
mkdir -p gpurun_out
i=0
for args in "$@"; do
  i=$((i+1))
  if [ "$N" = "1" ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+i)) bench.py"; fi
  $cmd --gpus $N $args --no-cpu-baseline > gpurun_out/${tag}_n${N}_$i.json 2> gpurun_out/${tag}_n${N}_$i.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_n${N}_$i.json").read().strip().splitlines()[-1]); k=d["kernels"]
    print("N=$N", "$args", "|", round(d["value"],2), "it/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"],2), "mfrg", round(d["mfrg_matvecs_per_sec_e2e"],1), d["state_sha1"], {x:round(k[x]["ms_per_step"],3) for x in k})
except Exception as e:
    print("N=$N $args failed:", e); print(open("gpurun_out/${tag}_n${N}_$i.err").read()[-1500:])
PY
done
