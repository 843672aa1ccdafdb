#!/bin/bash
# round 2, GPU call 25: ncu capture of the thin-lens retry kernel as it stands (compact tables, speculative pixel indices, merged normalisation)
tag=r02y
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:thin_persistent -s 2 -c 1 -o gpurun_out/${tag}_ncu_thin \
    python bench.py --workload config3 --spp 32 --steps 1 --warmup 2 --no-cpu --no-e2e --census-rays 0 > gpurun_out/${tag}_ncu_thin.log 2>&1
ls -la gpurun_out/${tag}*
