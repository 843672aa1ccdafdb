#!/bin/bash
# round 2 evidence on ONE B200 (what profiles/r02_* is made from; tools/profile_extracts.py turns the scratch files into
# the committed summaries): bench lines of every BASELINE configuration at N = 1, the launch list of the headline command,
# ncu --set full captures of the main kernel, the exact re-run kernel and the thin-lens kernel, compute-sanitizer on the
# kernels that are new this round.
tag=${1:-r02h}
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/${tag}_bench_headline.json 2> gpurun_out/${tag}_bench_headline.err
head -c 300 gpurun_out/${tag}_bench_headline.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_headline_reference.json 2> gpurun_out/${tag}_bench_headline_reference.err
timeout 300 python bench.py --workload config1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_config1.json 2> gpurun_out/${tag}_bench_config1.err
timeout 300 python bench.py --workload config2 --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_config2.json 2> gpurun_out/${tag}_bench_config2.err
timeout 300 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_config3.json 2> gpurun_out/${tag}_bench_config3.err
timeout 400 python bench.py --workload config4 --steps 3 --warmup 1 --no-cpu > gpurun_out/${tag}_bench_config4.json 2> gpurun_out/${tag}_bench_config4.err
for lens in double_gauss_f2.0.dat fisheye_muller_f4.0.dat mori_f2.8.dat petzval_f1.25.dat petzval_f1.6.dat telephoto_f5.0.dat tessar_f2.8.dat triplet_f2.5.dat; do
  timeout 400 python bench.py --workload config5:$lens --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/${tag}_bench_config5_${lens%.dat}.json 2> gpurun_out/${tag}_bench_config5_${lens%.dat}.err
done
for f in gpurun_out/${tag}_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d.get("parity_census") or {}
    print(d["config"]["workload"][:44], round(d["value"]), "Mrays/s", round(d["ms_per_step"],2), "ms frac", d.get("roofline") and round(d["roofline"]["frac"],4), "e2e", d.get("e2e") and round(d["e2e"]["value"]), "census flips", c.get("flips"), "of", c.get("rays"))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
# launch list of the headline command (cold-cache, serialised: compare shares)
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/${tag}_launches_headline.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --census-rays 0 > gpurun_out/${tag}_launches_headline.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/${tag}_launches_config3.csv python bench.py --workload config3 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/${tag}_launches_config3.log 2>&1
# full captures on a whole-film 33 M-ray batch
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kolb_pool2 -s 2 -c 1 -o gpurun_out/${tag}_ncu_pool2 \
    python bench.py --spp 8 --steps 1 --warmup 2 --no-cpu --no-e2e --census-rays 0 > gpurun_out/${tag}_ncu_pool2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kolb_exact_persistent -s 2 -c 1 -o gpurun_out/${tag}_ncu_rerun \
    python bench.py --spp 8 --steps 1 --warmup 2 --no-cpu --no-e2e --census-rays 0 > gpurun_out/${tag}_ncu_rerun.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:thin_persistent -s 2 -c 1 -o gpurun_out/${tag}_ncu_thin \
    python bench.py --workload config3 --spp 32 --steps 1 --warmup 2 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_thin.log 2>&1
# compute-sanitizer on this round's new kernels (job runner, census, consumer, LUT boxes, differentials, synthesis)
ZOICB_TEST_MAX_SAMPLES=4194304 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -x -q -m gpu tests/test_gpu_jobs.py \
    -k "streamed_job or census or differentials or rearm or lut_boxes" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" | head -8 > gpurun_out/${tag}_sanitizer_jobs.txt
cat gpurun_out/${tag}_sanitizer_jobs.txt
ls gpurun_out | grep ${tag} | wc -l
