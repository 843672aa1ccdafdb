#!/bin/bash
# round 2, GPU call 30: resident CTAs of the thin-lens retry kernel as it stands (compact tables: 32 registers without spills at 7 / 8 CTAs)
tag=r02ad
mkdir -p gpurun_out
rm -f gpurun_out/${tag}_ab.txt
for v in cta6 cta7 cta8 cta5 cta6 cta8; do
  ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/$v timeout 120 python bench.py --workload config3 --spp 32 --steps 5 --warmup 3 --no-cpu --no-e2e --census-rays 0 2>>gpurun_out/${tag}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3 spp32 $v', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab.txt
done
ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/cta8 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "thin" 2>&1 | tail -1 >> gpurun_out/${tag}_ab.txt
for v in cta6 cta8; do
  ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/$v timeout 200 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/${tag}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3 full $v', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab.txt
done
tail -3 gpurun_out/${tag}.err
cat gpurun_out/${tag}_ab.txt
