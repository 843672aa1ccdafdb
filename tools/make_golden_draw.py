"""Generate tests/golden/draw_rays_*.{txt,npz} FROM THE -D_DRAW BUILD OF THE UNMODIFIED REFERENCE (oracle/_ref).

Run in the build container (needs /root/reference):   python tools/make_golden_draw.py
The draw build writes ./draw.zoic: the lens header, then the (z, y) path of every attempt of ONE sample in 100 000
(src/zoic.cpp:1758-1764: the call that finds dd.counter == 100000 is drawn).  Each case feeds K * 100000 + 1 samples through
camera_create_ray, keeps the file and the K samples that were drawn (with their global indices, i.e. retry streams),
so that zoicb_write_draw_file can be asked for exactly those samples on a machine without the reference.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # the camera of the reference's own src/draw.zoic fixture; LUT sampling
    "dg_f28_focus23": dict(lensDataPath="double_gauss_f2.0.dat", focalLength=5.0, fStop=2.8, focalDistance=23.0),
    # no LUT: many attempts die at the rear rim / the stop, so partial paths are recorded
    "dg_nolut": dict(lensDataPath="double_gauss_f2.0.dat", focalLength=5.0, fStop=2.0, kolbSamplingLUT=0),
    # 12 elements, strong rejection, a zero-weight sample among the drawn ones
    "fisheye": dict(lensDataPath="fisheye_muller_f4.0.dat", focalLength=1.0, fStop=4.0),
}
K, SEED, FIRST = 24, 0xD7A3, 77_000_000

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
from oracle import ref
from zoic_b200.workloads import lens_path
kw = dict(%(kw)r); kw["lensDataPath"] = lens_path(kw["lensDataPath"])
cam = ref.RefCamera(draw=True, **kw)
s = np.load(%(samples)r)
cam.generate(s, seed=%(seed)d, first_index=%(first)d)
cam.close()
"""


def main():
    from oracle import ref
    assert ref.available(draw=True), "build oracle/_ref first (make -C oracle ref)"
    for name, kw in CASES.items():
        n = K * 100000 + 1
        rng = np.random.default_rng(sum(map(ord, name)))
        s = np.stack([rng.uniform(-1, 1, n), rng.uniform(-2 / 3, 2 / 3, n), rng.random(n), rng.random(n)], 1).astype(np.float32)
        tmp = os.path.join("/tmp", "zoic_draw_" + name)
        os.makedirs(tmp, exist_ok=True)
        np.save(os.path.join(tmp, "s.npy"), s)
        code = CHILD % dict(root=ROOT, kw=kw, samples=os.path.join(tmp, "s.npy"), seed=SEED, first=FIRST)
        subprocess.run([sys.executable, "-c", code], cwd=tmp, check=True, capture_output=True)
        text = open(os.path.join(tmp, "draw.zoic")).read()
        # the call that finds dd.counter == 100000 is drawn and resets the counter to 0 (it is 1 again when the call
        # returns): calls 100000, 200000, 300000, ...
        drawn = np.array([(k + 1) * 100000 for k in range(K)], np.int64)
        open(os.path.join(OUT, "draw_rays_%s.txt" % name), "w").write(text)
        np.savez_compressed(os.path.join(OUT, "draw_rays_%s.npz" % name), samples=s[drawn], index=drawn + FIRST,
                            seed=np.int64(SEED))
        rays = text[text.index("RAYS{") + 5:text.rindex("}")].split()
        print(name, "drawn samples", len(drawn), "numbers in RAYS{}", len(rays), "bytes", len(text))


if __name__ == "__main__":
    main()
