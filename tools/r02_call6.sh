#!/bin/bash
# round 2, GPU call 6 (one GPU): where does the time of the high-rejection cameras go?  launch lists (main / exact re-run)
tag=r02f
mkdir -p gpurun_out
for wl in config4 config5:telephoto_f5.0.dat config5:tessar_f2.8.dat headline; do
  n=${wl#config5:}; n=${n%.dat}
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_${n}.csv python bench.py --workload $wl --spp 16 --steps 2 --warmup 1 --no-cpu --no-e2e --census-rays 0 > gpurun_out/${tag}_launches_${n}.log 2>&1
  python - <<PY
import csv
from collections import defaultdict
rows=[r for r in csv.reader(open("gpurun_out/${tag}_launches_${n}.csv", errors="replace")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
acc=defaultdict(list)
for r in rows[1:]:
    try: acc[r[ik][:48]].append(float(r[iv].replace(",","")))
    except: pass
print("$wl")
for k,v in acc.items(): print("   %-50s n=%2d mean %10.1f us" % (k,len(v),sum(v)/len(v)/1e3))
PY
done
