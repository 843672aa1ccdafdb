"""The oracle's CPU restatement (oracle/zoic_port.cpp) against every fixture that pins the reference:

* tests/golden/rays_*.npz + golden.json -- outputs of the compiled, UNMODIFIED reference (tools/make_golden.py),
* tests/golden/draw_zoic_header.txt     -- lines 1-10 of the reference's own src/draw.zoic, its one known-answer test,
* the argument-evaluation order of the reference's two-draw call sites (golden.json: draw_order_pins).

Runs without a GPU and without the reference tree.
"""
import os

import numpy as np
import pytest

from zutil import (GOLDEN, bits_equal, draw_zoic_header, golden_case, golden_index, golden_names, setup_log_values)


@pytest.mark.parametrize("name", golden_names())
def test_port_reproduces_reference_rays_bit_for_bit(port, name):
    kw, image, meta, s, o_ref, d_ref = golden_case(name)
    cam = port.PortCamera(image=image, **kw)
    o, d, st = cam.generate(s, seed=meta["seed"], first_index=meta["first_index"])
    assert bits_equal(o, o_ref), name
    assert bits_equal(d, d_ref), name
    assert st["attempts"] == meta["stats"]["attempts"]
    assert (o[:, 3] == 0).sum() == meta["stats"]["vignetted"]
    cam.close()


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith("kolb")])
def test_port_setup_matches_reference_log(port, name):
    """The numbers node_update prints (focal lengths, aperture radius, image distance ...), as %12.8f strings."""
    kw, image, meta, *_ = golden_case(name)
    cam = port.PortCamera(image=image, **kw)
    got = setup_log_values(cam.constants(), kw["fStop"])
    for key, val in meta["setup_log"].items():
        assert got[key] == val, (name, key, got[key], val)
    cam.close()


def test_port_reproduces_draw_zoic_known_answer(port):
    """reference src/draw.zoic:1-10 -- Double Gauss, focalLength 5.0, fStop 2.8, focalDistance 23.0."""
    from zoic_b200.workloads import lens_path
    want = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "draw_zoic_header.txt"))]
    cam = port.PortCamera(lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.8, focalDistance=23.0)
    got = draw_zoic_header(cam.constants(), 23.0)
    assert got == want
    cam.close()


def test_port_draw_order_matches_reference(port):
    """Of the two xor128() calls in one argument list, the reference build feeds the FIRST draw to the SECOND
    parameter.  Hand-made stream states, rays that needed at least one retry, every two-draw call site."""
    idx = golden_index()
    for name, rows in idx["draw_order_pins"].items():
        kw, image, *_ = golden_case(name)
        cam = port.PortCamera(image=image, **kw)
        for row in rows:
            o, d = cam.generate_one(row["sample"], row["state"])
            assert bits_equal(o, np.array(row["origin_w"], np.float32)), name
            assert bits_equal(d, np.array(row["dir_tries"], np.float32)), name
            assert d[3] > 0
        cam.close()


def test_port_is_thread_count_invariant(port):
    kw, image, meta, s, o_ref, d_ref = golden_case("kolb_fisheye")
    cam = port.PortCamera(image=image, **kw)
    big = np.tile(s, (40, 1))
    o1, d1, st1 = cam.generate(big, seed=3, first_index=10, nthreads=1)
    o8, d8, st8 = cam.generate(big, seed=3, first_index=10, nthreads=8)
    assert bits_equal(o1, o8) and bits_equal(d1, d8) and st1 == st8
    cam.close()


def test_synthetic_samples_definition(port):
    """Pixel-major / spp-minor sample grid with 24-bit uniforms: ranges and pixel mapping."""
    W, H, spp = 64, 48, 4
    s = port.synth_samples(W, H, spp, 0x200C, 0, W * H * spp)
    px = np.arange(W * H * spp) // spp % W
    py = np.arange(W * H * spp) // spp // W
    u0 = (s[:, 0] + 1) * W / 2 - px
    u1 = (1 - s[:, 1] * W / H) * H / 2 - py
    assert (u0 > -1e-3).all() and (u0 < 1 + 1e-3).all() and (u1 > -1e-3).all() and (u1 < 1 + 1e-3).all()
    assert (s[:, 2:] >= 0).all() and (s[:, 2:] < 1).all()
    assert abs(s[:, 1]).max() <= H / W + 1e-6
    # any slice can be regenerated independently
    part = port.synth_samples(W, H, spp, 0x200C, 1000, 500)
    assert bits_equal(part, s[1000:1500])


def test_epilogue_contract_statement_is_an_affine_map(port):
    """oracle.port.transform_rays (the CPU statement of zoicb_transform_rays): agrees with a float64 matrix product
    to fp32 rounding, passes weight / tries through, and composes with its inverse to the identity."""
    rng = np.random.default_rng(5)
    rays = rng.normal(size=(5000, 8)).astype(np.float32)
    m = rng.normal(size=(3, 4)).astype(np.float32)
    out = port.transform_rays(rays, m)
    M = m.astype(np.float64)
    o = rays[:, :3].astype(np.float64) @ M[:, :3].T + M[:, 3]
    d = rays[:, 4:7].astype(np.float64) @ M[:, :3].T
    assert np.abs(out[:, :3] - o).max() < 2e-6 * max(1.0, np.abs(o).max())
    assert np.abs(out[:, 4:7] - d).max() < 2e-6 * max(1.0, np.abs(d).max())
    assert np.array_equal(out[:, 3], rays[:, 3]) and np.array_equal(out[:, 7], rays[:, 7])
    inv = np.linalg.inv(np.vstack([M, [0, 0, 0, 1]]))[:3].astype(np.float32)
    back = port.transform_rays(out, inv)
    assert np.abs(back[:, :3] - rays[:, :3]).max() < 1e-3 and np.abs(back[:, 4:7] - rays[:, 4:7]).max() < 1e-3
