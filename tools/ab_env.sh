# A/B of environment settings on the GPU box: tools/ab_env.sh "<bench args>" "VAR=1 VAR2=x" "VAR=0" ...
mkdir -p gpurun_out
args="$1"; shift
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python bench.py $args --no-cpu-baseline > gpurun_out/abenv_$i.json 2> gpurun_out/abenv_$i.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/abenv_$i.json').read().strip().splitlines()[-1]); k=d['kernels']; print('$envs', '|', round(d['value'],2), round(d['e2e']['value'],2), d['state_sha1'], {x:round(k[x]['ms_per_step'],3) for x in ('K2','L_K2','sde_L','column_K2','swave')})"
done
