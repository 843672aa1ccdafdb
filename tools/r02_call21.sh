#!/bin/bash
# round 2, GPU call 21: ncu captures of the two thin-lens retry kernels (shipped schedule / prepared blocks), config 3 at 32 spp
tag=r02u
mkdir -p gpurun_out
ZOICB_THIN_PREP=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:thin_persistent -s 2 -c 1 -o gpurun_out/${tag}_ncu_thin0 \
    python bench.py --workload config3 --spp 32 --steps 1 --warmup 2 --no-cpu --no-e2e --census-rays 0 > gpurun_out/${tag}_ncu_thin0.log 2>&1
ZOICB_THIN_PREP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:thin_prepared -s 2 -c 1 -o gpurun_out/${tag}_ncu_thin1 \
    python bench.py --workload config3 --spp 32 --steps 1 --warmup 2 --no-cpu --no-e2e --census-rays 0 > gpurun_out/${tag}_ncu_thin1.log 2>&1
ls -la gpurun_out/${tag}*
