#!/bin/bash
# round 2, GPU call 31: thin-lens compact kernel at 8 CTAs with 10-byte row tables (32 KB carve-out, 224 KB L1); thin evidence refreshed
tag=r02ae
mkdir -p gpurun_out
rm -f gpurun_out/${tag}_ab.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jobs.py -x -q -m gpu -k "thin or config3 or bokeh or streamed or small or planar" 2>&1 | tail -2 >> gpurun_out/${tag}_ab.txt
for rep in 1 2; do
  timeout 120 python bench.py --workload config3 --spp 32 --steps 5 --warmup 3 --no-cpu --no-e2e --census-rays 0 2>>gpurun_out/${tag}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3 spp32 rows10', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab.txt
done
timeout 300 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_config3.json 2>>gpurun_out/${tag}.err
python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench_config3.json').read().strip().splitlines()[-1]); print('config3 full rows10', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms', 'e2e', round(d['e2e']['value']))" >> gpurun_out/${tag}_ab.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/${tag}_launches_config3.csv python bench.py --workload config3 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/${tag}_launches_config3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:thin_persistent -s 2 -c 1 -o gpurun_out/${tag}_ncu_thin \
    python bench.py --workload config3 --spp 32 --steps 1 --warmup 2 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_thin.log 2>&1
tail -3 gpurun_out/${tag}.err
cat gpurun_out/${tag}_ab.txt
