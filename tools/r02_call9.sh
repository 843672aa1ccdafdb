#!/bin/bash
tag=r02i
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jobs.py -x -q -m gpu -k "not full_size" ) > gpurun_out/${tag}_pytest.log 2>&1
tail -4 gpurun_out/${tag}_pytest.log
line() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms', 'frac', round(d['roofline']['frac'],4))
except Exception as e: print('$1 FAILED', e)
"; }
B="--no-cpu --no-e2e --census-rays 0"
timeout 300 python bench.py --steps 5 --warmup 3 $B 2>>gpurun_out/${tag}.err | line "headline resident" >> gpurun_out/${tag}_ab.txt
timeout 300 python bench.py --workload config4 --steps 3 --warmup 1 $B 2>>gpurun_out/${tag}.err | line "config4 streamed" >> gpurun_out/${tag}_ab.txt
timeout 300 python bench.py --workload config4 --samples 2123366400 --steps 3 --warmup 1 $B 2>>gpurun_out/${tag}.err | tee gpurun_out/${tag}_bench_config4_resident_half.json | line "config4 resident 2.1G" >> gpurun_out/${tag}_ab.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 $B > /dev/null 2>&1
grep -E "kolb_pool2|kolb_exact" gpurun_out/${tag}_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120 | tail -4
cat gpurun_out/${tag}_ab.txt
