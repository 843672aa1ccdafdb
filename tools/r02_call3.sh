#!/bin/bash
# round 2, GPU call 3 (one GPU): why is the streamed job slower than the resident batch?  serial vs pipelined, with and
# without the consumer; thin-lens carve-out.
tag=r02c
mkdir -p gpurun_out
line() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms', 'frac', round(d['roofline']['frac'],4))
except Exception as e: print('$1 FAILED', e)
"; }
B="--steps 5 --warmup 3 --no-cpu --no-e2e --census-rays 0"
timeout 300 python bench.py --workload headline --stream $B 2>>gpurun_out/${tag}.err | line "headline streamed pipelined" >> gpurun_out/${tag}_ab.txt
timeout 300 python bench.py --workload headline --stream --serial $B 2>>gpurun_out/${tag}.err | line "headline streamed serial" >> gpurun_out/${tag}_ab.txt
ZOICB_JOB_NO_CONSUME=1 timeout 300 python bench.py --workload headline --stream $B 2>>gpurun_out/${tag}.err | line "headline streamed pipelined no-consume" >> gpurun_out/${tag}_ab.txt
ZOICB_JOB_NO_CONSUME=1 timeout 300 python bench.py --workload headline --stream --serial $B 2>>gpurun_out/${tag}.err | line "headline streamed serial no-consume" >> gpurun_out/${tag}_ab.txt
timeout 300 python bench.py --workload config3 $B 2>>gpurun_out/${tag}.err | tee gpurun_out/${tag}_bench_config3.json | line "config3 full (carve-out auto, col guide 2w)" >> gpurun_out/${tag}_ab.txt
for c in 0 14 28 44; do
  ZOICB_THIN_CARVEOUT=$c timeout 300 python bench.py --workload config3 --spp 32 $B 2>>gpurun_out/${tag}.err | line "config3 spp32 carveout $c%" >> gpurun_out/${tag}_ab.txt
done
cat gpurun_out/${tag}_ab.txt
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "thin or bokeh or image" ) > gpurun_out/${tag}_pytest_parity.log 2>&1
tail -4 gpurun_out/${tag}_pytest_parity.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches_streamed.csv python bench.py --workload headline --stream --steps 1 --warmup 1 --no-cpu --no-e2e --census-rays 0 > gpurun_out/${tag}_launches_streamed.log 2>&1
tail -3 gpurun_out/${tag}.err
