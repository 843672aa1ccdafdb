#!/bin/bash
# ncu --set full of the pool kernel on one config-5 camera: tools/prof_lens.sh <lens file> <out name> <ZOICB_POOL value>
lens=$1; out=$2; pool=${3:-0}
cat > /tmp/_prof_lens.py <<PY
import os, sys, torch
sys.path.insert(0, os.getcwd())
from zoic_b200 import ZoicCamera, workloads
wl = workloads.config5("$lens"); wl.spp = 1
cam = ZoicCamera(**wl.params)
s = cam.synth_samples(wl.W, wl.H, wl.spp, wl.seed, 0, wl.n)
rays = torch.empty((wl.n, 8), dtype=torch.float32, device="cuda")
for _ in range(3): cam.create_rays(s, seed=wl.seed, out=rays)
torch.cuda.synchronize()
PY
ZOICB_POOL=$pool ncu --set full --clock-control none --import-source on -k regex:kolb_pool -s 2 -c 1 -o gpurun_out/$out python /tmp/_prof_lens.py > gpurun_out/$out.log 2>&1
tail -1 gpurun_out/$out.log | cut -c1-200
