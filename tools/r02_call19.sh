#!/bin/bash
# round 2, GPU call 19: the march's surface constants from the constant bank (LDC / LDCU) instead of shared memory (LDS)
tag=r02s
mkdir -p gpurun_out
for v in ldsfull ldcfull ldsfull ldcfull; do
  ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/$v timeout 300 python bench.py --spp 32 --steps 5 --warmup 3 --no-cpu --no-e2e --census-rays 0 2>>gpurun_out/${tag}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headline spp32 $v', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab.txt
done
ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/ldcfull timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "guarded_kolb_all_lenses" 2>&1 | tail -1 >> gpurun_out/${tag}_ab.txt
cat gpurun_out/${tag}_ab.txt
