mkdir -p gpurun_out
for tw in 0 16 11 8; do
  FDGA_CONV_TW=$tw python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s5_tw$tw.json 2>gpurun_out/s5_tw$tw.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/s5_tw$tw.json').read().strip().splitlines()[-1]); k=d['kernels']; print('TW',$tw, round(d['value'],1), round(d['e2e']['value'],1), {x:round(k[x]['ms_per_step'],3) for x in ('K2','L_K2','sde_L','column_K2')})"
done
