#!/bin/bash
# round 2, GPU call 7 (one GPU): the whole GPU suite after the resumable re-run, then bench lines.
tag=r02g
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -6 gpurun_out/${tag}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
line() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms', 'frac', round(d['roofline']['frac'],4), 'reruns', d['stats']['exact_reruns'])
except Exception as e: print('$1 FAILED', e)
"; }
B="--no-cpu --no-e2e --census-rays 0"
timeout 300 python bench.py --steps 5 --warmup 3 $B 2>>gpurun_out/${tag}.err | tee gpurun_out/${tag}_bench_headline.json | line "headline resident" >> gpurun_out/${tag}_ab.txt
timeout 300 python bench.py --stream --steps 5 --warmup 3 $B 2>>gpurun_out/${tag}.err | line "headline streamed" >> gpurun_out/${tag}_ab.txt
timeout 300 python bench.py --workload config4 --steps 3 --warmup 1 $B 2>>gpurun_out/${tag}.err | tee gpurun_out/${tag}_bench_config4.json | line "config4" >> gpurun_out/${tag}_ab.txt
timeout 300 python bench.py --workload config2 --steps 5 --warmup 3 $B 2>>gpurun_out/${tag}.err | tee gpurun_out/${tag}_bench_config2.json | line "config2" >> gpurun_out/${tag}_ab.txt
timeout 300 python bench.py --workload config5:tessar_f2.8.dat --steps 2 --warmup 1 $B 2>>gpurun_out/${tag}.err | line "config5 tessar" >> gpurun_out/${tag}_ab.txt
cat gpurun_out/${tag}_ab.txt; tail -3 gpurun_out/${tag}.err
