#!/bin/bash
# thin lens + optical vignetting + bokeh image (config3, 32 spp = 265 M samples): resident CTAs per SM of the persistent kernel
for v in ../lib; do
  ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/$v python bench.py --workload config1 --spp 256 --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v thin-no-OV 530M', round(d['value']), 'Mrays/s', round(d['ms_per_step'],2), 'ms', 'hbm frac', round(d['roofline']['frac'],4))
except Exception as e: print('$v FAILED', e)
"
done
for v in ../lib thin4t thin5 thin6 thin6t; do
  ZOICB_LIBDIR=$PWD/zoic_b200/lib_variants/$v python bench.py --workload config3 --spp 32 --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v config3', round(d['value']), 'Mrays/s', round(d['ms_per_step'],2), 'ms', 'hbm frac', round(d['roofline']['frac'],4))
except Exception as e: print('$v config3 FAILED', e)
"
done
