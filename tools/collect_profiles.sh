#!/bin/bash
# round evidence on one B200: bench lines of the BASELINE configs, the ncu launch list of the headline command, one
# ncu --set full capture of the main kernel and of the exact re-run kernel.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r01}
python bench.py --workload config4 --steps 3 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_config4.json 2> gpurun_out/${tag}_bench_config4.err
python bench.py --workload config3 --steps 3 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_config3.json 2> gpurun_out/${tag}_bench_config3.err
python bench.py --workload config2 --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_config2.json 2> gpurun_out/${tag}_bench_config2.err
python bench.py --workload config1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_config1.json 2> gpurun_out/${tag}_bench_config1.err
# launch list of the headline command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/${tag}_launches_headline.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/${tag}_launches_headline.log 2>&1
# full captures on a whole-film 33 M-ray batch
ncu --set full --clock-control none --import-source on -k regex:kolb_pool2 -s 2 -c 1 -o gpurun_out/${tag}_ncu_pool2 \
    python bench.py --spp 4 --steps 1 --warmup 2 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_pool2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kolb_exact_persistent -s 2 -c 1 -o gpurun_out/${tag}_ncu_rerun \
    python bench.py --spp 4 --steps 1 --warmup 2 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_rerun.log 2>&1
for f in gpurun_out/${tag}_bench_config*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["config"]["workload"][:40], round(d["value"]), "Mrays/s", round(d["ms_per_step"],2), "ms frac", round(d["roofline"]["frac"],4), "e2e", d["e2e"] and round(d["e2e"]["value"]))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
