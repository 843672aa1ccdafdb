#!/bin/bash
tag=r02l
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jobs.py -x -q -m gpu -k "not full_size" ) > gpurun_out/${tag}_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/${tag}_pytest.log | tail -2
line() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms', 'frac', round(d['roofline']['frac'],4))
except Exception as e: print('$1 FAILED', e)
"; }
B="--no-cpu --no-e2e --census-rays 0"
timeout 300 python bench.py --steps 5 --warmup 3 $B 2>>gpurun_out/${tag}.err | line "headline resident (8 CTAs, 68-byte slots)" >> gpurun_out/${tag}_ab.txt
timeout 300 python bench.py --spp 32 --steps 5 --warmup 3 $B 2>>gpurun_out/${tag}.err | line "headline spp32" >> gpurun_out/${tag}_ab.txt
for d in zoic_b200/lib_variants/*/; do v=$(basename $d)
  ZOICB_LIBDIR=$PWD/$d timeout 300 python bench.py --spp 32 --steps 5 --warmup 3 $B 2>>gpurun_out/${tag}.err | line "headline spp32 $v" >> gpurun_out/${tag}_ab.txt
done
timeout 300 python bench.py --workload config4 --samples 2123366400 --steps 3 --warmup 1 $B 2>>gpurun_out/${tag}.err | line "config4 resident 2.1G" >> gpurun_out/${tag}_ab.txt
timeout 300 python bench.py --workload config5:tessar_f2.8.dat --samples 2123366400 --steps 3 --warmup 1 $B 2>>gpurun_out/${tag}.err | line "tessar resident 2.1G" >> gpurun_out/${tag}_ab.txt
timeout 300 python bench.py --workload config4 --steps 3 --warmup 1 $B 2>>gpurun_out/${tag}.err | line "config4 streamed" >> gpurun_out/${tag}_ab.txt
cat gpurun_out/${tag}_ab.txt; tail -2 gpurun_out/${tag}.err
