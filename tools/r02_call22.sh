#!/bin/bash
# round 2, GPU call 22: thin-lens retry kernel with byte-wide column tables + the rows' final CDF values in shared memory
# (ZOICB_THIN_COMPACT=1, default) against the 16-bit tables (=0); column guide resolution with the narrow tables
tag=r02v
mkdir -p gpurun_out
rm -f gpurun_out/${tag}_ab.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jobs.py -x -q -m gpu -k "thin or config3 or bokeh or streamed or small" 2>&1 | tail -3 >> gpurun_out/${tag}_ab.txt
ZOICB_THIN_COMPACT=0 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "thin" 2>&1 | tail -1 >> gpurun_out/${tag}_ab.txt
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --workload config3 --spp 32 --steps 5 --warmup 3 --no-cpu --no-e2e --census-rays 0 2>>gpurun_out/${tag}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3 spp32 $label', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab.txt
}
for rep in 1 2; do
  run compact=0 ZOICB_THIN_COMPACT=0
  run compact=1 ZOICB_THIN_COMPACT=1
done
run compact=1,colguide=2^8 ZOICB_THIN_COMPACT=1 ZOICB_GUIDE_COL_LOG2=8
run compact=1,colguide=2^10 ZOICB_THIN_COMPACT=1 ZOICB_GUIDE_COL_LOG2=10
run compact=0,colguide=2^8 ZOICB_THIN_COMPACT=0 ZOICB_GUIDE_COL_LOG2=8
run compact=1,carve=8 ZOICB_THIN_COMPACT=1 ZOICB_THIN_CARVEOUT=8
for v in 0 1; do
  ZOICB_THIN_COMPACT=$v timeout 300 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/${tag}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3 full compact=$v', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab.txt
done
tail -5 gpurun_out/${tag}.err
cat gpurun_out/${tag}_ab.txt
