#!/bin/bash
# round 2, GPU call 16: thin-lens kernel with a prepared block of 32 samples (dense set-up + seeding, adopted through shuffles)
tag=r02o
mkdir -p gpurun_out
for v in base6 prep6 prep5 prep4; do
  ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/$v timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "guarded_thin_lens_is_bit_exact" 2>&1 | tail -1 >> gpurun_out/${tag}_ab.txt
  ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/$v timeout 300 python bench.py --workload config3 --spp 32 --steps 5 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/${tag}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3 spp32 $v', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab.txt
done
cat gpurun_out/${tag}_ab.txt
