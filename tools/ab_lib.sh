# A/B several builds of libfdga on the GPU box: tools/ab_lib.sh "<bench args>" libfdga.so libfdga_exp.so ...
mkdir -p gpurun_out
args="$1"; shift
for lib in "$@"; do
  FDGA_LIB_PATH=$PWD/fddgasolver.jl_b200/$lib python bench.py $args --no-cpu-baseline > gpurun_out/ab_$lib.json 2> gpurun_out/ab_$lib.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/ab_$lib.json').read().strip().splitlines()[-1]); k=d['kernels']; print('$lib', '$args', round(d['value'],2), round(d['e2e']['value'],2), d['state_sha1'], {x:round(k[x]['ms_per_step'],3) for x in ('K2','L_K2','sde_L','column_K2')})"
done
