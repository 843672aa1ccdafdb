#!/bin/bash
# times every prebuilt variant under zoic_b200/lib_variants/ (tools/build_variants.py) on the GPU box:
# headline camera at 32 spp (265 M rays) and the fisheye camera (config4) at 8 spp (265 M rays)
for d in zoic_b200/lib_variants/${1:-*}/; do   # optional argument: a glob of variant names
  v=$(basename $d)
  for wl in headline config4; do
    spp=32; [ $wl = config4 ] && spp=8
    ZOICB_LIBDIR=$PWD/$d python bench.py --workload $wl --spp $spp --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v $wl', round(d['value']), 'Mrays/s', round(d['ms_per_step'],2), 'ms', 'frac', round(d['roofline']['frac'],4))
except Exception as e: print('$v $wl FAILED', e)
"
  done
done
