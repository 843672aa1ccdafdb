#!/bin/bash
# scalar (ZOICB_POOL=1) against packed (ZOICB_POOL=2) and packed + rim pre-test (ZOICB_POOL=3) pool kernel on the fisheye (config4, 8 spp) and on three
# config-5 cameras with very different rejection rates; 265 M rays each
run() { # name, python expression building the workload
python - "$1" "$2" <<'PY'
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
from zoic_b200 import ZoicCamera, workloads
name, lens = sys.argv[1], sys.argv[2]
wl = workloads.config4() if lens == "config4" else workloads.config5(lens)
wl.spp = 8
cam = ZoicCamera(**wl.params)
n = wl.n
s = cam.synth_samples(wl.W, wl.H, wl.spp, wl.seed, 0, n)
rays = torch.empty((n, 8), dtype=torch.float32, device="cuda")
for _ in range(2): cam.create_rays(s, seed=wl.seed, out=rays)
torch.cuda.synchronize(); cam.reset_stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): cam.create_rays(s, seed=wl.seed, out=rays)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
st = cam.stats(); c = cam.constants()
print("%-8s %-26s %8.0f Mrays/s  %7.2f ms  attempts/ray %.2f visits/ray %.2f zero-weight %.3f split %d inner %d" % (
    name, lens, n / ms / 1e3, ms, st["attempts"] / st["rays"], st["element_visits"] / st["rays"], st["vignetted"] / st["rays"],
    c["guardedSplit"], c["guardedInnerRetry"]))
PY
}
for lens in ${LENSES:-config4 telephoto_f5.0.dat petzval_f1.25.dat tessar_f2.8.dat double_gauss_f2.0.dat}; do
  ZOICB_POOL=1 run scalar $lens
  ZOICB_POOL=2 run packed $lens
  ZOICB_POOL=3 run pretest $lens
  ZOICB_POOL=0 run auto $lens
done
