#!/usr/bin/env python
"""GUARDED-vs-EXACT census on RANDOM cameras: node parameters drawn over their documented ranges (src/zoic.mtd; the sweep
of tests/test_gpu_parity.py::test_random_cameras_sweep, at scale), every lens table, with and without the exit-pupil LUT;
2^28 rays of a 4K x 64 spp frame per camera, every record compared on the device (zoicb_run_job's census).
usage (GPU box): python tools/census_random.py [cameras] > profiles/r02_census_random.txt"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zoic_b200 import ZoicCamera
from zoic_b200.workloads import LENSES, lens_path

ncam = int(sys.argv[1]) if len(sys.argv) > 1 else 48
rng = np.random.default_rng(20261017)
lenses = sorted(LENSES)
n = 1 << 28
tot = dict(rays=0, flips=0, bad=0)
print("# lens, focalLength, fStop, focalDistance, sensorWidth, LUT, exposure | rays, flips, out of tolerance, max rel origin, max dir, re-run %, attempts/ray")
for k in range(ncam):
    lens = lenses[k % len(lenses)]
    native = 1.0 if "fisheye" in lens else 5.0
    kw = dict(lensModel=1, lensDataPath=lens_path(lens), focalLength=float(native * rng.uniform(0.6, 1.8)),
              fStop=float(rng.uniform(1.2, 11.0)), focalDistance=float(rng.uniform(25.0, 800.0)),
              kolbSamplingLUT=int(rng.random() < 0.8), exposureControl=float(rng.uniform(-1, 1)),
              sensorWidth=float(rng.choice([3.6, 2.4, 1.8])))
    cam = ZoicCamera(**kw)
    r = cam.run_job(3840, 2160, 8, 1000 + k, 2000 + k, int(rng.integers(0, 1 << 30)), n, census=True)
    st = r["stats"]
    print("%-26s %6.3f %6.3f %7.2f %4.1f %d %6.3f | %10d %6d %6d %9.2e %9.2e %6.3f %6.2f" % (
        lens, kw["focalLength"], kw["fStop"], kw["focalDistance"], kw["sensorWidth"], kw["kolbSamplingLUT"], kw["exposureControl"],
        r["census_rays"], r["census_flips"], r["census_out_of_tol"], r["census_max_rel_origin"], r["census_max_dir"],
        100.0 * st["exact_reruns"] / n, st["attempts"] / n), flush=True)
    tot["rays"] += r["census_rays"]; tot["flips"] += r["census_flips"]; tot["bad"] += r["census_out_of_tol"]
    cam.close()
print("# total: %d rays, %d flips, %d out of tolerance" % (tot["rays"], tot["flips"], tot["bad"]))
