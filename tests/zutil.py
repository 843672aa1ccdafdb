"""Shared helpers for the parity tests."""
import numpy as np


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
