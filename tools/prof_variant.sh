#!/bin/bash
# ncu --set full capture of the pool kernel of one prebuilt variant: tools/prof_variant.sh <variant dir name> <out name> [workload] [spp] [kernel regex]   (variant "../lib" = the main library)
v=$1; out=$2; wl=${3:-headline}; spp=${4:-4}; kern=${5:-kolb_pool}
ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/$v ncu --set full --clock-control none --import-source on -k regex:$kern -s 2 -c 1 \
  -o gpurun_out/$out python bench.py --workload $wl --spp $spp --steps 1 --warmup 2 --no-cpu --no-e2e > gpurun_out/$out.log 2>&1
tail -1 gpurun_out/$out.log | cut -c1-200
