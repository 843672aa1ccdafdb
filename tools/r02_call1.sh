#!/bin/bash
# round 2, GPU call 1 (one GPU): new job runner / census / LUT tests, regression of the parity suite with the exact-cell
# guide tables, first bench lines of every workload, thin-lens guide-resolution A/B, ncu of the thin-lens kernel.
tag=r02a
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv >> gpurun_out/${tag}_topo.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_jobs.py -x -q -m gpu -k "not full_size" -s ) > gpurun_out/${tag}_pytest_jobs.log 2>&1
tail -5 gpurun_out/${tag}_pytest_jobs.log
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_capi_library.py -x -q -m gpu ) > gpurun_out/${tag}_pytest_parity.log 2>&1
tail -5 gpurun_out/${tag}_pytest_parity.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
# thin lens + image: guide resolution and counting threshold
for g in 6 7 8 9 10; do
  echo "col guide log2 = $g" >> gpurun_out/${tag}_ab_thin.txt
  ZOICB_GUIDE_COL_LOG2=$g timeout 300 python bench.py --workload config3 --spp 32 --steps 5 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/${tag}_ab_thin.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab_thin.txt 2>&1
done
for d in zoic_b200/lib_variants/*/; do
  v=$(basename $d)
  echo "variant $v" >> gpurun_out/${tag}_ab_thin.txt
  ZOICB_LIBDIR=$PWD/$d timeout 300 python bench.py --workload config3 --spp 32 --steps 5 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/${tag}_ab_thin.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')" >> gpurun_out/${tag}_ab_thin.txt 2>&1
done
cat gpurun_out/${tag}_ab_thin.txt
( time timeout 600 python bench.py ) > gpurun_out/${tag}_bench_headline.json 2> gpurun_out/${tag}_bench_headline.err
head -c 600 gpurun_out/${tag}_bench_headline.json; echo
timeout 300 python bench.py --workload config3 --no-cpu --no-e2e > gpurun_out/${tag}_bench_config3.json 2> gpurun_out/${tag}_bench_config3.err
head -c 300 gpurun_out/${tag}_bench_config3.json; echo
timeout 300 python bench.py --workload config4 --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/${tag}_bench_config4.json 2> gpurun_out/${tag}_bench_config4.err
head -c 300 gpurun_out/${tag}_bench_config4.json; echo
timeout 300 python bench.py --workload headline --stream --steps 3 --warmup 1 --no-cpu --no-e2e --census-rays 0 > gpurun_out/${tag}_bench_headline_streamed.json 2> gpurun_out/${tag}_bench_headline_streamed.err
head -c 300 gpurun_out/${tag}_bench_headline_streamed.json; echo
( time timeout 900 python -m pytest tests/test_gpu_jobs.py -x -q -m gpu -k "full_size" -s ) > gpurun_out/${tag}_pytest_fullsize.log 2>&1
grep -E "rays in|passed|failed|Error|error" gpurun_out/${tag}_pytest_fullsize.log | head -20
timeout 300 ncu --set full --clock-control none --import-source on -k regex:thin_persistent -s 2 -c 1 -o gpurun_out/${tag}_ncu_thin \
    python bench.py --workload config3 --spp 4 --steps 1 --warmup 2 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_thin.log 2>&1
ls -la gpurun_out | tail -20
