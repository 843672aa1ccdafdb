#!/bin/bash
# round 2, GPU call 29: ncu capture of the ring variant of the thin-lens kernel
tag=r02ac
mkdir -p gpurun_out
ZOICB_THIN_RING=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:thin_ring -s 2 -c 1 -o gpurun_out/${tag}_ncu_ring \
    python bench.py --workload config3 --spp 32 --steps 1 --warmup 2 --no-cpu --no-e2e --census-rays 0 > gpurun_out/${tag}_ncu_ring.log 2>&1
ls -la gpurun_out/${tag}*
