"""Generate tests/golden/*.npz|json FROM THE COMPILED, UNMODIFIED REFERENCE (oracle/_ref).

Run in the build container (needs /root/reference):   python tools/make_golden.py
The vectors pin the oracle restatement and the CUDA path on machines where the reference does not exist.
Every case stores its inputs (camera parameters, samples, seed, first index) and the reference's outputs.
"""
import json
import os
import re
import subprocess
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402
from zoic_b200.synth import hex_bokeh_image  # noqa: E402
from zoic_b200.workloads import LENSES, lens_path  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def samples(n, seed):
    rng = np.random.default_rng(seed)
    s = np.stack([rng.uniform(-1, 1, n), rng.uniform(-2 / 3, 2 / 3, n), rng.random(n), rng.random(n)], 1).astype(np.float32)
    # edge cases: image corners, axis points, lens-square corners and edges
    edge = np.array([[1, 2 / 3, 0, 0], [-1, -2 / 3, 1 - 2 ** -24, 1 - 2 ** -24], [1, -2 / 3, 0, 1 - 2 ** -24],
                     [0.5, 0, 0.25, 0.75], [0, 0.5, 0.75, 0.25], [1e-3, 1e-3, 0.5, 0.25], [-0.3, 0.2, 0.5, 0.75],
                     [0.125 / 1.8, 0, 0.1, 0.9]], np.float32)
    s[:len(edge)] = edge
    return s


def kolb(lens, **kw):
    fnum, focal = LENSES[lens]
    d = dict(lensModel=1, lensDataPath=lens, focalLength=focal, fStop=fnum)
    d.update(kw)
    return d


CASES = {
    "thin_plain": (dict(lensModel=0, focalLength=3.5, fStop=2.8), None),
    "thin_nodof_exposure": (dict(lensModel=0, focalLength=3.5, fStop=2.8, useDof=0, exposureControl=0.5), None),
    "thin_ov": (dict(lensModel=0, focalLength=3.5, fStop=2.8, opticalVignettingDistance=2.0, opticalVignettingRadius=1.0,
                     exposureControl=-1.25), None),
    "thin_ov_harsh": (dict(lensModel=0, focalLength=2.0, fStop=1.4, opticalVignettingDistance=4.0,
                           opticalVignettingRadius=0.6), None),
    "thin_ov_hex33": (dict(lensModel=0, focalLength=3.5, fStop=2.8, opticalVignettingDistance=2.0, useImage=1), 33),
    "kolb_dg_lut": (kolb("double_gauss_f2.0.dat"), None),
    "kolb_dg_f28_focus23": (kolb("double_gauss_f2.0.dat", fStop=2.8, focalDistance=23.0), None),
    "kolb_dg_nolut": (kolb("double_gauss_f2.0.dat", kolbSamplingLUT=0, exposureControl=1.5), None),
    "kolb_dg_lut_hex33": (kolb("double_gauss_f2.0.dat", useImage=1), 33),
    "kolb_fisheye": (kolb("fisheye_muller_f4.0.dat"), None),
    "kolb_petzval16_nostop": (kolb("petzval_f1.6.dat"), None),
    "kolb_telephoto": (kolb("telephoto_f5.0.dat"), None),
    "kolb_tessar": (kolb("tessar_f2.8.dat"), None),
    "kolb_triplet": (kolb("triplet_f2.5.dat"), None),
    "kolb_mori": (kolb("mori_f2.8.dat"), None),
    "kolb_petzval125": (kolb("petzval_f1.25.dat"), None),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    meta = {}
    n, seed, first = 384, 0x5EED, 1_000_003
    if len(sys.argv) == 1:
        # one fresh process per case: the reference leaves `apertureElement` uninitialised for lens tables
        # without a stop (src/zoic.cpp:532), and only a fresh heap reads as the 0 the rulings assume
        for name in list(CASES) + ["--pins"]:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), name])
        parts = [json.load(open(os.path.join(OUT, "_part_%s.json" % k))) for k in list(CASES) + ["--pins"]]
        cases = {}
        for p in parts[:-1]:
            cases.update(p)
        json.dump({"cases": cases, "draw_order_pins": parts[-1]}, open(os.path.join(OUT, "golden.json"), "w"), indent=1)
        for k in list(CASES) + ["--pins"]:
            os.remove(os.path.join(OUT, "_part_%s.json" % k))
        ref_draw = "/root/reference/src/draw.zoic"  # the one externally authored known-answer test: lines 1-10
        if os.path.exists(ref_draw):
            with open(ref_draw) as f:
                head = [next(f) for _ in range(10)]
            open(os.path.join(OUT, "draw_zoic_header.txt"), "w").writelines(head)
        return
    only = sys.argv[1]
    for name, (params, img) in CASES.items():
        if name != only:
            continue
        image = hex_bokeh_image(img) if img else None
        kw = dict(params)
        if "lensDataPath" in kw:
            kw["lensDataPath"] = lens_path(kw["lensDataPath"])
        cam = ref.RefCamera(image=image, **kw)
        s = samples(n, zlib.crc32(name.encode()) % 1000)
        o, d, st = cam.generate(s, seed=seed, first_index=first)
        log = cam.log
        cam.close()
        np.savez_compressed(os.path.join(OUT, "rays_%s.npz" % name), samples=s, origin_w=o, dir_tries=d)
        setup = {}
        for key in ("Principle Plane distance", "Focal point distance", "Raytraced Focal Length", "Focal length ratio",
                    "Adj. PP distance", "Adj. Focal point distance", "Adj. Raytraced Focal Length",
                    "User aperture radius", "Image distance", "Aperture distance", "Aperture is lens element number"):
            m = re.search(r"\[ZOIC\] " + re.escape(key) + r"[^\n]*?\s(-?[0-9.]+)\n", log)
            if m:
                setup[key] = m.group(1)
        meta[name] = {"params": params, "hex_image": img, "n": n, "seed": seed, "first_index": first, "stats": st,
                      "setup_log": setup}
        print(name, st, len(setup))
        json.dump(meta, open(os.path.join(OUT, "_part_%s.json" % name), "w"))
    if only != "--pins":
        return
    # argument-evaluation order of the two-draw call sites: one retried ray per call site with a hand-made stream
    pins = {}
    for name, state in (("thin_ov", [1, 2, 3, 4]), ("kolb_dg_lut", [5, 6, 7, 8]), ("kolb_dg_nolut", [9, 10, 11, 12]),
                        ("thin_ov_hex33", [13, 14, 15, 16]), ("kolb_dg_lut_hex33", [17, 18, 19, 20])):
        params, img = CASES[name]
        image = hex_bokeh_image(img) if img else None
        kw = dict(params)
        if "lensDataPath" in kw:
            kw["lensDataPath"] = lens_path(kw["lensDataPath"])
        cam = ref.RefCamera(image=image, **kw)
        rows = []
        rng = np.random.default_rng(7)
        while len(rows) < 6:  # keep rays that needed at least one retry
            s = np.array([rng.uniform(-1, 1), rng.uniform(-0.6, 0.6), rng.random(), rng.random()], np.float32)
            st = (np.array(state, np.uint64) * 2654435761 + len(rows) * 97 + int(rng.integers(1 << 30))) % (1 << 32)
            o, d, dv = cam.generate_one(s, st.astype(np.uint32))
            if d[3] > 0:
                rows.append({"sample": s.tolist(), "state": [int(x) for x in st], "origin_w": o.tolist(),
                             "dir_tries": d.tolist(), "derivs": dv.tolist()})
        cam.close()
        pins[name] = rows
    json.dump(pins, open(os.path.join(OUT, "_part_--pins.json"), "w"))


if __name__ == "__main__":
    main()
