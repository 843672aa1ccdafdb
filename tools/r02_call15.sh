#!/bin/bash
tag=r02n
mkdir -p gpurun_out
for g in 8 10 12 14; do
  ZOICB_GUIDE_ROW_LOG2=$g timeout 300 python bench.py --workload config3 --spp 32 --steps 5 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/${tag}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3 spp32 row guide 2^$g', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab.txt
done
cat gpurun_out/${tag}_ab.txt
