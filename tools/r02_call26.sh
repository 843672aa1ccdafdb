#!/bin/bash
# round 2, GPU call 26: thin-lens compact path, straight-line row search for brackets of at most two entries (row1) against the counting loop (row0)
tag=r02z
mkdir -p gpurun_out
rm -f gpurun_out/${tag}_ab.txt
ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/row1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jobs.py -x -q -m gpu -k "thin or config3 or bokeh or streamed or small or normalisation" 2>&1 | tail -3 >> gpurun_out/${tag}_ab.txt
for rep in 1 2; do
for v in row0 row1; do
  ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/$v timeout 300 python bench.py --workload config3 --spp 32 --steps 5 --warmup 3 --no-cpu --no-e2e --census-rays 0 2>>gpurun_out/${tag}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3 spp32 $v', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab.txt
done
done
for v in row0 row1; do
  ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/$v timeout 300 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/${tag}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3 full $v', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab.txt
done
tail -5 gpurun_out/${tag}.err
cat gpurun_out/${tag}_ab.txt
