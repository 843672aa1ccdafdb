#!/bin/bash
# round 2, GPU call 33: column guide resolution again, with the shipped thin-lens kernel (64 warps per SM)
tag=r02aj
mkdir -p gpurun_out
rm -f gpurun_out/${tag}_ab.txt
for g in 9 8 10 9 8; do
  ZOICB_GUIDE_COL_LOG2=$g timeout 120 python bench.py --workload config3 --spp 32 --steps 5 --warmup 3 --no-cpu --no-e2e --census-rays 0 2>>gpurun_out/${tag}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3 spp32 colguide=2^$g', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab.txt
done
cat gpurun_out/${tag}_ab.txt
