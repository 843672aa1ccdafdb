#!/bin/bash
# A/B of the pool kernel's CTA shape on the GPU box: rebuilds libzoicb with different -D flags and runs the bench.
set -e
for cfg in "3 8" "4 7" "4 6" "2 12"; do
  set -- $cfg
  ZOICB_NVCC_FLAGS="-DZOICB_POOL_CTAS=$1 -DZOICB_POOL_WARPS=$2" python zoic_b200/build.py --force > /dev/null 2>&1
  python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e | python -c "import json,sys; d=json.load(sys.stdin); print('ctas=$1 warps=$2 headline', round(d['value']), round(d['ms_per_step'],1))"
  python bench.py --workload config4 --steps 2 --warmup 2 --no-cpu --no-e2e | python -c "import json,sys; d=json.load(sys.stdin); print('ctas=$1 warps=$2 config4 ', round(d['value']), round(d['ms_per_step'],1))"
done
