#!/bin/bash
# round 2, GPU call 10: how much does residency buy the pool kernel?  5 / 6 / 7 CTAs of 3 warps per SM (headline, 32 spp);
# margin head-room census (tools/census_scales.py) for the headline camera and the fisheye.
tag=r02j
mkdir -p gpurun_out
for v in ctas5 ctas6 ctas7; do
  ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/$v timeout 300 python bench.py --workload headline --spp 32 --steps 5 --warmup 3 --no-cpu --no-e2e --census-rays 0 2>>gpurun_out/${tag}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab.txt
done
cat gpurun_out/${tag}_ab.txt
timeout 600 python tools/census_scales.py headline config4 config5:tessar_f2.8.dat config5:telephoto_f5.0.dat > gpurun_out/${tag}_census_scales.txt 2>>gpurun_out/${tag}.err
cat gpurun_out/${tag}_census_scales.txt
