#!/bin/bash
# round 2, GPU call 2 (one GPU): differentials / LUT quirk tests, image-sampling rewrite (regression + speed), tile-size and
# chunk-size A/B of the streamed job.
tag=r02b
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_jobs.py -x -q -m gpu -k "not full_size" -s ) > gpurun_out/${tag}_pytest_jobs.log 2>&1
tail -4 gpurun_out/${tag}_pytest_jobs.log
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu ) > gpurun_out/${tag}_pytest_parity.log 2>&1
tail -4 gpurun_out/${tag}_pytest_parity.log
line() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms', 'wall', d['run'].get('wall_ms_per_step'), 'frac', round(d['roofline']['frac'],4))
except Exception as e: print('$1 FAILED', e)
"; }
B="--steps 5 --warmup 3 --no-cpu --no-e2e --census-rays 0"
timeout 300 python bench.py --workload config3 $B 2>>gpurun_out/${tag}.err | tee gpurun_out/${tag}_bench_config3.json | line "config3 full" >> gpurun_out/${tag}_ab.txt
for g in 8 9; do
  ZOICB_GUIDE_COL_LOG2=$g timeout 300 python bench.py --workload config3 --spp 32 $B 2>>gpurun_out/${tag}.err | line "config3 spp32 colguide$g" >> gpurun_out/${tag}_ab.txt
done
for v in count2 count4 thin5; do
  ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/$v timeout 300 python bench.py --workload config3 --spp 32 $B 2>>gpurun_out/${tag}.err | line "config3 spp32 $v" >> gpurun_out/${tag}_ab.txt
done
for t in 26 27 28 29; do
  timeout 300 python bench.py --workload headline --stream --tile-log2 $t $B 2>>gpurun_out/${tag}.err | line "headline streamed tile2^$t" >> gpurun_out/${tag}_ab.txt
done
timeout 300 python bench.py --workload headline --spp 32 $B 2>>gpurun_out/${tag}.err | line "headline spp32 resident" >> gpurun_out/${tag}_ab.txt
for v in chunk512 chunk1024 chunk4096; do
  ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/$v timeout 300 python bench.py --workload headline --spp 32 $B 2>>gpurun_out/${tag}.err | line "headline spp32 resident $v" >> gpurun_out/${tag}_ab.txt
  ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/$v timeout 300 python bench.py --workload headline --stream --tile-log2 27 $B 2>>gpurun_out/${tag}.err | line "headline streamed tile2^27 $v" >> gpurun_out/${tag}_ab.txt
done
timeout 300 python bench.py --workload config4 --steps 3 --warmup 1 --no-cpu --no-e2e --census-rays 0 --tile-log2 28 2>>gpurun_out/${tag}.err | line "config4 tile2^28" >> gpurun_out/${tag}_ab.txt
cat gpurun_out/${tag}_ab.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:thin_persistent -s 2 -c 1 -o gpurun_out/${tag}_ncu_thin \
    python bench.py --workload config3 --spp 16 --steps 1 --warmup 2 --no-cpu --no-e2e --census-rays 0 > gpurun_out/${tag}_ncu_thin.log 2>&1
tail -3 gpurun_out/${tag}.err
