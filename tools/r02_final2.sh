#!/bin/bash
# round 2, last GPU call: launch list of the headline command on the final tree (traffic.json), then what the driver runs
tag=${1:-r02af}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/${tag}_launches_headline.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --census-rays 0 > gpurun_out/${tag}_launches_headline.log 2>&1
bash tools/round_check.sh ${tag}
