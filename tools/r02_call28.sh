#!/bin/bash
# round 2, GPU call 28: thin-lens retry kernel with a producer warp per CTA feeding the others through a shared-memory ring
# (ZOICB_THIN_RING=1) against the shipped schedule (=0)
tag=r02ab
mkdir -p gpurun_out
rm -f gpurun_out/${tag}_ab.txt
ZOICB_THIN_RING=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "guarded_thin_lens_is_bit_exact" 2>&1 | tail -3 >> gpurun_out/${tag}_ab.txt
for rep in 1 2; do
for v in 0 1; do
  ZOICB_THIN_RING=$v timeout 120 python bench.py --workload config3 --spp 32 --steps 5 --warmup 3 --no-cpu --no-e2e --census-rays 0 2>>gpurun_out/${tag}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3 spp32 ring=$v', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab.txt
done
done
ZOICB_THIN_RING=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jobs.py -x -q -m gpu -k "thin or config3 or bokeh or streamed or small" 2>&1 | tail -3 >> gpurun_out/${tag}_ab.txt
for v in 0 1; do
  ZOICB_THIN_RING=$v timeout 200 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/${tag}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3 full ring=$v', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab.txt
done
tail -5 gpurun_out/${tag}.err
cat gpurun_out/${tag}_ab.txt
