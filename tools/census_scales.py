#!/usr/bin/env python
"""Head-room of the GUARDED mode's decision margins, measured on whole jobs: for every camera the first 2.1 G samples of
its BASELINE frame are generated with the margins scaled by 1, 1/2, 1/4, 1/8 and 0 (zoicb_set_guard_scale) and every record
is compared with the EXACT mode's on the device (zoicb_run_job's census).  Prints one line per (camera, scale):
flips = records whose weight or tries differ, re-runs = samples the fast path handed to the exact kernel.
usage (GPU box): python tools/census_scales.py [workload ...] > profiles/r02_census_scales.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zoic_b200 import ZoicCamera, workloads

names = sys.argv[1:] or ["headline", "config4"] + workloads.CONFIG5
print("# camera, margin scale, rays, flips, out of tolerance (1e-5), max rel origin, max dir, exact re-runs")
for name in names:
    wl = workloads.BY_NAME[name]()
    cam = ZoicCamera(**wl.params)
    n = min(wl.n, 2_123_366_400)
    for scale in (1.0, 0.5, 0.25, 0.125, 0.0):
        cam.set_guard_scale(scale)
        r = cam.run_job(*wl.synth_args(), wl.seed, 0, n, census=True)
        print("%-34s %5.3f %11d %9d %9d %9.2e %9.2e %10d" % (name, scale, r["census_rays"], r["census_flips"], r["census_out_of_tol"],
              r["census_max_rel_origin"], r["census_max_dir"], r["stats"]["exact_reruns"]), flush=True)
    cam.close()
