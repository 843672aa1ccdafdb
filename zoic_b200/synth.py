"""Synthetic workloads: the procedural hexagonal bokeh image.

Everything here is deterministic integer / IEEE arithmetic so that every machine regenerates
byte-identical inputs instead of shipping them.
"""
import numpy as np


def hex_bokeh_image(size=255):
    """size x size x 3 float32 image of a regular hexagon inscribed in the unit disk.

    Inside the hexagon the value is 1 - 0.25*r plus a per-pixel 1e-4 hash dither (so no two luminances
    inside the aperture tie); outside it is exactly 0.  Odd `size` keeps a centre pixel, which the
    reference's centring arithmetic assumes (reference src/zoic.cpp:440).
    """
    c = (size - 1) / 2.0
    y, x = np.mgrid[0:size, 0:size]
    u = (x - c) / c
    v = (y - c) / c
    h = np.float64(0.8660254037844386)
    inside = np.maximum(np.abs(v), np.abs(u) * h + np.abs(v) * 0.5) <= h
    r = np.sqrt(u * u + v * v)
    idx = (y * size + x).astype(np.uint64)
    hashed = ((idx * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)) >> np.uint64(8)
    dither = hashed.astype(np.float64) / 16777216.0
    val = np.where(inside, 1.0 - 0.25 * r + 1e-4 * dither, 0.0).astype(np.float32)
    return np.ascontiguousarray(np.repeat(val[:, :, None], 3, axis=2))
