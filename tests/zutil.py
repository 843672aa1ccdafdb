"""Shared helpers for the parity tests."""
import json
import math
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def random_samples(n, seed=1, aspect=1.5):
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(-1, 1, n), rng.uniform(-1 / aspect, 1 / aspect, n), rng.random(n), rng.random(n)],
                    1).astype(np.float32)


def bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def compare_rays(o, d, o_ref, d_ref, tol=1e-5):
    """The north-star tolerance, per vector:
    |d_origin| <= tol * max(|origin|, 1 cm), |d_dir| <= tol, weight and tries exact;
    origin/dir only where weight != 0 (zero-weight rays carry the half-traced state of the last failed attempt).
    Returns a dict of counts."""
    o, d, o_ref, d_ref = (np.asarray(x, np.float64) for x in (o, d, o_ref, d_ref))
    flips = (o[:, 3] != o_ref[:, 3]) | (d[:, 3] != d_ref[:, 3])
    live = (~flips) & (o_ref[:, 3] != 0)
    do = np.linalg.norm(o[live, :3] - o_ref[live, :3], axis=1)
    dd = np.linalg.norm(d[live, :3] - d_ref[live, :3], axis=1)
    scale = np.maximum(np.linalg.norm(o_ref[live, :3], axis=1), 1.0)
    bad = (do > tol * scale) | (dd > tol) | ~np.isfinite(do) | ~np.isfinite(dd)
    return {"n": len(o), "path_flips": int(flips.sum()), "out_of_tol": int(bad.sum()),
            "max_origin_err": float((do / scale).max()) if live.any() else 0.0,
            "max_dir_err": float(dd.max()) if live.any() else 0.0}


# ---------------------------------------------------------------------------------------------------
# golden vectors (generated from the compiled, unmodified reference by tools/make_golden.py)
# ---------------------------------------------------------------------------------------------------
def golden_index():
    return json.load(open(os.path.join(GOLDEN, "golden.json")))


def golden_case(name):
    from zoic_b200.synth import hex_bokeh_image
    from zoic_b200.workloads import lens_path
    meta = golden_index()["cases"][name]
    data = np.load(os.path.join(GOLDEN, "rays_%s.npz" % name))
    kw = dict(meta["params"])
    if "lensDataPath" in kw:
        kw["lensDataPath"] = lens_path(kw["lensDataPath"])
    image = hex_bokeh_image(meta["hex_image"]) if meta["hex_image"] else None
    return kw, image, meta, data["samples"], data["origin_w"], data["dir_tries"]


def golden_names():
    return sorted(golden_index()["cases"])


# ---------------------------------------------------------------------------------------------------
# the reference's draw.zoic header (reference src/zoic.cpp:1240-1293, written with setprecision(10))
# ---------------------------------------------------------------------------------------------------
def draw_zoic_header(constants, focal_distance):
    """Format derived camera state the way the reference's writeToFile() does (lines 1-10 of draw.zoic)."""
    f = lambda v: "%.10f" % float(v)
    lenses = constants["lenses"]  # columns: curvature, thickness, ior, aperture, center
    deg = float(np.float32(180) / np.float32(3.14159265358979323846))  # 180 / AI_PI is a float division
    out = ["LENSMODEL{KOLB}"]
    parts = []
    for cur, _th, _ior, ap, cen in lenses:
        ang = math.asin((float(ap) * 0.5) / float(cur)) * deg
        parts += [f(-cen), f(-cur), f(ang)]
    out.append("LENSES{" + " ".join(parts) + " }")
    out.append("IOR{" + " ".join(f(r[2]) for r in lenses) + " }")
    out.append("APERTUREELEMENT{%d}" % constants["apertureElement"])
    out.append("APERTUREDISTANCE{" + f(-constants["apertureDistance"]) + "}")
    out.append("APERTURE{" + f(constants["userApertureRadius"]) + "}")
    out.append("APERTUREMAX{" + f(max(float(r[3]) for r in lenses)) + "}")
    out.append("FOCUSDISTANCE{" + f(-np.float32(focal_distance)) + "}")
    out.append("IMAGEDISTANCE{" + f(-constants["originShift"]) + "}")
    out.append("SENSORHEIGHT{" + f(1.7) + "}")
    return out


def setup_log_values(c, fstop):
    """The values the reference prints with AiMsgInfo("%12.8f") during node_update, from derived constants.
    The aperture radius is printed BEFORE the clamp to the lens table's aperture (src/zoic.cpp:1664-1672)."""
    g = lambda v: "%.8f" % float(v)
    unclamped = np.float32(float(c["tracedFocalLength1"]) / (2.0 * float(np.float32(fstop))))
    assert np.float32(c["userApertureRadius"]) in (unclamped, np.float32(c["lenses"][c["apertureElement"]][3]))
    return {"Principle Plane distance": g(c["principalPlane0"]), "Focal point distance": g(c["focalPoint0"]),
            "Raytraced Focal Length": g(c["tracedFocalLength0"]), "Focal length ratio": g(c["focalLengthRatio"]),
            "Adj. PP distance": g(c["principalPlane1"]), "Adj. Focal point distance": g(c["focalPoint1"]),
            "Adj. Raytraced Focal Length": g(c["tracedFocalLength1"]),
            "User aperture radius": g(unclamped), "Image distance": g(c["originShift"]),
            "Aperture distance": g(c["apertureDistance"]),
            "Aperture is lens element number": "%d" % c["apertureElement"]}


def product_constants_flat(c):
    """zoic_b200.host_setup()/ZoicCamera.constants() dict -> the flat names oracle.port uses."""
    out = dict(c)
    for k in ("tracedFocalLength", "principalPlane", "focalPoint"):
        out[k + "0"], out[k + "1"] = c[k][0], c[k][1]
    return out
