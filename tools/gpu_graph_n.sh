#!/bin/bash
# multi-rank CUDA-graph replay (NCCL collectives captured): tools/gpu_graph_n.sh N [extra bench args]
N=$1; shift
for G in 0 1; do
  echo "== FDGA_GRAPH_MULTIRANK=$G"
  FDGA_GRAPH_MULTIRANK=$G timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$G bench.py --gpus $N --steps 200 --warmup 10 "$@" 2>gpurun_out/graph_n${N}_g$G.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['n_gpus'], 'mfrg', d.get('mfrg_matvecs_per_sec_e2e'), 'kry', d.get('mfrg_dqgmres_iterations_per_sec_device_resident'), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'sha', d['state_sha1'], d['config'].get('issue'))
"
done
