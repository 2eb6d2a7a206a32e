#!/bin/bash
# device-resident K2 timing per (MWB_K2_VARIANT, MWB_K2_FLAGS) pair; usage: gpu_variants2.sh "v:f v:f ..."
mkdir -p gpurun_out
for vf in ${1:-"1:3 0:3 2:3"}; do
  v=${vf%%:*}; f=${vf#*:}
  MWB_K2_VARIANT=$v MWB_K2_FLAGS=$f MWB_DEBUG=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu 2> gpurun_out/var_${v}_${f}.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant=$v flags=$f value=%.0f k2=%.4f k1=%.4f e2e=%.0f' % (d['value'], d['roofline']['kernel_avg_ms'], d['roofline']['k1_avg_ms'], d['e2e']['value']))"
  grep "K2 variant" gpurun_out/var_${v}_${f}.err | head -1
done
