"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on identical inputs.

EXACT mode must reproduce the oracle BIT FOR BIT (origin, dir, weight, tries and the counters); the
north-star tolerance (1e-5 relative per vector) is therefore met with zero path flips.
"""
import numpy as np
import pytest

from zutil import bits_equal, compare_rays, random_samples

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _run_gpu(cam, s, seed, first_index=0):
    t = torch.from_numpy(s).cuda()
    cam.reset_stats()
    rays = cam.create_rays(t, seed=seed, first_index=first_index)
    torch.cuda.synchronize()
    r = rays.cpu().numpy()
    return np.ascontiguousarray(r[:, :4]), np.ascontiguousarray(r[:, 4:]), cam.stats()


def _check_exact(kw, port, n=200_000, seed=11, image=None):
    from zoic_b200 import ZoicCamera, MODE_EXACT
    cam = ZoicCamera(image=image, mode=MODE_EXACT, **kw)
    ref = port.PortCamera(image=image, **kw)
    s = random_samples(n, seed=seed)
    o, d, st = _run_gpu(cam, s, seed=seed, first_index=12345)
    o2, d2, st2 = ref.generate(s, seed=seed, first_index=12345, nthreads=8)
    res = compare_rays(o, d, o2, d2)
    assert res["path_flips"] == 0 and res["out_of_tol"] == 0, res
    # zero-weight rays hold the half-traced state of the last attempt: compare their bits too (same arithmetic)
    assert bits_equal(o, o2) and bits_equal(d, d2), res
    assert st["rays"] == n
    assert st["attempts"] == st2["attempts"]
    assert st["element_visits"] == st2["element_visits"]
    assert st["total_internal_reflection"] == st2["tir"]
    if kw.get("lensModel", 1) == 1 or kw.get("useDof", 1):
        assert st["success"] == st2["success"] and st["vignetted"] == st2["vignetted"]
    cam.close()
    ref.close()


def test_thin_lens_plain(port):
    _check_exact(dict(lensModel=0, focalLength=3.5, fStop=2.8), port)


def test_thin_lens_no_dof(port):
    _check_exact(dict(lensModel=0, focalLength=3.5, fStop=2.8, useDof=0, exposureControl=0.5), port)


def test_thin_lens_optical_vignetting(port):
    _check_exact(dict(lensModel=0, focalLength=3.5, fStop=2.8, opticalVignettingDistance=2.0,
                      opticalVignettingRadius=1.0, exposureControl=-1.25), port)


def test_thin_lens_hex_bokeh(port):
    from zoic_b200.synth import hex_bokeh_image
    _check_exact(dict(lensModel=0, focalLength=3.5, fStop=2.8, opticalVignettingDistance=2.0, useImage=1), port,
                 image=hex_bokeh_image(255))


def test_thin_lens_nonsquare_even_bokeh(port):
    from zoic_b200.synth import hex_bokeh_image
    img = hex_bokeh_image(33)[:, :20].copy()
    _check_exact(dict(lensModel=0, focalLength=3.5, fStop=2.8, opticalVignettingDistance=2.0, useImage=1), port,
                 image=img, n=50_000)


@pytest.mark.parametrize("lens", ["double_gauss_f2.0.dat", "fisheye_muller_f4.0.dat", "petzval_f1.6.dat",
                                  "telephoto_f5.0.dat", "tessar_f2.8.dat", "triplet_f2.5.dat", "mori_f2.8.dat",
                                  "petzval_f1.25.dat"])
def test_kolb_lut_all_lenses(port, lens):
    from zoic_b200.workloads import LENSES, lens_path
    fnum, focal = LENSES[lens]
    _check_exact(dict(lensModel=1, lensDataPath=lens_path(lens), focalLength=focal, fStop=fnum), port, n=60_000)


def test_kolb_no_lut(port):
    from zoic_b200.workloads import lens_path
    _check_exact(dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.0,
                      kolbSamplingLUT=0, exposureControl=1.5), port, n=60_000)


def test_kolb_lut_hex_bokeh(port):
    from zoic_b200.synth import hex_bokeh_image
    from zoic_b200.workloads import lens_path
    _check_exact(dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.0,
                      useImage=1), port, n=60_000, image=hex_bokeh_image(255))


@pytest.mark.parametrize("nch", [1, 2])
def test_bokeh_image_with_fewer_than_three_channels(port, nch):
    """A 1- or 2-channel image is what it is in the reference: accepted, invalid (src/zoic.cpp:135-137), every aperture
    sample the lens centre (:420-425) -- pinned against the compiled reference in tests/test_oracle_vs_reference.py."""
    from zoic_b200.workloads import lens_path
    img = np.random.default_rng(nch).random((9, 7, nch)).astype(np.float32)
    _check_exact(dict(lensModel=0, focalLength=3.5, fStop=2.8, opticalVignettingDistance=2.0, useImage=1), port, n=50_000, image=img)
    _check_exact(dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.0, useImage=1),
                 port, n=50_000, image=img)
    _check_guarded(dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.0, useImage=1),
                   port, n=100_000, image=img)


def test_synth_samples_match_oracle(port):
    from zoic_b200 import ZoicCamera
    cam = ZoicCamera(lensModel=0, focalLength=3.5, fStop=2.8)
    for (W, H, spp, first, n) in [(1920, 1080, 1, 0, 100_000), (3840, 2160, 256, 2_000_000_000, 100_000),
                                  (7680, 4320, 1024, 33_973_000_000, 65_536)]:
        g = cam.synth_samples(W, H, spp, 0x200C, first, n).cpu().numpy()
        c = port.synth_samples(W, H, spp, 0x200C, first, n)
        assert bits_equal(g, c)
    cam.close()


def test_batch_boundaries_do_not_matter(port):
    """Results depend on (seed, global sample index) only: splitting a batch changes nothing."""
    from zoic_b200 import ZoicCamera, MODE_EXACT
    from zoic_b200.workloads import lens_path
    cam = ZoicCamera(mode=MODE_EXACT, lensModel=1, lensDataPath=lens_path("fisheye_muller_f4.0.dat"),
                     focalLength=1.0, fStop=4.0)
    s = random_samples(50_000, seed=3)
    o, d, _ = _run_gpu(cam, s, seed=9, first_index=0)
    oa, da, _ = _run_gpu(cam, s[:20_001], seed=9, first_index=0)
    ob, db, _ = _run_gpu(cam, s[20_001:], seed=9, first_index=20_001)
    assert bits_equal(o, np.concatenate([oa, ob])) and bits_equal(d, np.concatenate([da, db]))
    cam.close()


def test_host_buffer_entry_point(port):
    from zoic_b200 import ZoicCamera, MODE_EXACT
    from zoic_b200.workloads import lens_path
    kw = dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.0)
    cam = ZoicCamera(mode=MODE_EXACT, **kw)
    ref = port.PortCamera(**kw)
    s = random_samples(300_000, seed=5)
    o2, d2, _ = ref.generate(s, seed=4, first_index=77, nthreads=8)
    r = cam.create_rays_host(s, seed=4, first_index=77)                # pageable numpy memory
    assert bits_equal(r[:, :4], o2) and bits_equal(r[:, 4:], d2)
    sp = torch.from_numpy(s).pin_memory()
    rp = torch.empty((len(s), 8), dtype=torch.float32).pin_memory()
    cam.create_rays_host(sp, seed=4, first_index=77, out=rp)           # pinned memory, direct DMA
    assert bits_equal(rp.numpy()[:, :4], o2) and bits_equal(rp.numpy()[:, 4:], d2)
    cam.close()
    ref.close()


def test_planar_host_output_is_lossless():
    """zoicb_generate_host_planar carries 25 bytes per ray over the host link instead of 32: the 32-byte records rebuilt
    from the planes equal zoicb_generate_host's bit for bit -- raytraced and thin lens (vignetted rays: weight 0, many
    tries), an exposure scale other than 1, pageable and pinned memory, more than one pipeline chunk."""
    from zoic_b200 import ZoicCamera, unpack_planes
    from zoic_b200.synth import hex_bokeh_image
    from zoic_b200.workloads import lens_path
    cams = [(None, dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.0, exposureControl=0.7)),
            (None, dict(lensModel=0, focalLength=2.0, fStop=1.4, opticalVignettingDistance=4.0, opticalVignettingRadius=0.6,
                        exposureControl=-0.5)),
            (hex_bokeh_image(33), dict(lensModel=0, focalLength=3.5, fStop=2.8, useImage=1, opticalVignettingDistance=2.0))]
    for image, kw in cams:
        cam = ZoicCamera(image=image, **kw)
        n = (1 << 21) * 2 + 12_345          # three chunks of the host pipeline, the last one ragged
        s = random_samples(n, seed=17)
        want = cam.create_rays_host(s, seed=9, first_index=3)
        planes, flags, w = cam.create_rays_host_planar(s, seed=9, first_index=3)            # pageable
        got = unpack_planes(planes, flags, w)
        assert bits_equal(got, want)
        assert (flags & 0x80).any() or kw["lensModel"] == 1 or image is not None   # the harsh thin lens has zero-weight rays
        sp = torch.from_numpy(s).pin_memory()
        pp = torch.empty((6, n), dtype=torch.float32).pin_memory()
        fp = torch.empty((n,), dtype=torch.uint8).pin_memory()
        _, _, w2 = cam.create_rays_host_planar(sp, seed=9, first_index=3, planes=pp, flags=fp)   # pinned: direct DMA
        assert w2 == w and bits_equal(unpack_planes(pp.numpy(), fp.numpy(), w2), want)
        cam.close()


def test_two_host_threads_share_a_camera_and_a_stream(port):
    """generate calls on one context from several host threads.  Two threads drive the SAME camera on the
    SAME (default) stream, guarded mode (counter reset + pool kernel + re-run kernel per call): every call's records must
    be those of a lone call -- the enqueue phases serialise on the context's lock, so no call sees another's cursor or
    queue -- and growing the workspace under way (a larger batch in between) must not pull a queue from under a caller."""
    import threading
    from zoic_b200 import ZoicCamera
    from zoic_b200.workloads import lens_path
    cam = ZoicCamera(lensModel=1, lensDataPath=lens_path("fisheye_muller_f4.0.dat"), focalLength=1.0, fStop=4.0)
    sizes = [70_000, 200_000, 90_000, 400_000]
    batches = [torch.from_numpy(random_samples(n, seed=40 + k)).cuda() for k, n in enumerate(sizes)]
    want = [cam.create_rays(b, seed=3, first_index=1000 * k).clone() for k, b in enumerate(batches)]
    torch.cuda.synchronize()
    errors = []

    def worker(tid):
        try:
            for rep in range(12):
                k = (rep + tid) % len(batches)
                out = cam.create_rays(batches[k], seed=3, first_index=1000 * k)
                torch.cuda.synchronize()
                if not torch.equal(out.view(torch.int32), want[k].view(torch.int32)):
                    errors.append((tid, rep, k))
        except Exception as exc:   # noqa: BLE001
            errors.append((tid, repr(exc)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    cam.close()


def test_empty_batch_and_errors():
    from zoic_b200 import ZoicCamera, capi
    cam = ZoicCamera(lensModel=0, focalLength=3.5, fStop=2.8)
    t = torch.empty((0, 4), dtype=torch.float32, device="cuda")
    r = cam.create_rays(t)
    assert r.shape == (0, 8)
    cam.close()
    with pytest.raises(capi.ZoicError) as e:
        ZoicCamera(lensModel=1, lensDataPath="/nonexistent/lens.dat")
    assert e.value.code == capi.ERR_LENS_FILE
    with pytest.raises(capi.ZoicError) as e:
        ZoicCamera(lensModel=0, useImage=1)
    assert e.value.code == capi.ERR_BOKEH_IMAGE


# ---------------------------------------------------------------------------------------------------
# GUARDED mode (the default): fused fast path + exact re-run of undecided samples.  Same decision sequence
# as the oracle (zero path flips: weight and tries identical everywhere, counters identical), values within
# the north-star tolerance: |d_origin| <= 1e-5 * max(|origin|, 1 cm), |d_dir| <= 1e-5 per ray.
# ---------------------------------------------------------------------------------------------------
def _check_guarded(kw, port, n=400_000, seed=21, image=None):
    from zoic_b200 import ZoicCamera, MODE_GUARDED
    cam = ZoicCamera(image=image, **kw)
    assert cam.mode == MODE_GUARDED  # the default
    ref = port.PortCamera(image=image, **kw)
    s = random_samples(n, seed=seed)
    o, d, st = _run_gpu(cam, s, seed=seed, first_index=999)
    o2, d2, st2 = ref.generate(s, seed=seed, first_index=999, nthreads=8)
    res = compare_rays(o, d, o2, d2, tol=1e-5)
    assert res["path_flips"] == 0 and res["out_of_tol"] == 0, res
    # zero-weight rays too must be well-formed records (film point + optical axis, or the exact path's last attempt)
    assert np.isfinite(o).all() and np.isfinite(d).all()
    assert st["rays"] == n and st["attempts"] == st2["attempts"] and st["element_visits"] == st2["element_visits"]
    assert st["success"] == st2["success"] and st["vignetted"] == st2["vignetted"]
    assert st["total_internal_reflection"] == st2["tir"]
    cam.close()
    ref.close()
    return res, st


@pytest.mark.parametrize("lens", ["double_gauss_f2.0.dat", "fisheye_muller_f4.0.dat", "petzval_f1.6.dat",
                                  "telephoto_f5.0.dat", "tessar_f2.8.dat", "triplet_f2.5.dat", "mori_f2.8.dat",
                                  "petzval_f1.25.dat"])
def test_guarded_kolb_all_lenses(port, lens):
    from zoic_b200.workloads import LENSES, lens_path
    fnum, focal = LENSES[lens]
    res, st = _check_guarded(dict(lensModel=1, lensDataPath=lens_path(lens), focalLength=focal, fStop=fnum), port)
    assert st["exact_reruns"] < 0.02 * st["rays"]


def test_guarded_kolb_no_lut_and_bokeh(port):
    from zoic_b200.synth import hex_bokeh_image
    from zoic_b200.workloads import lens_path
    _check_guarded(dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.0,
                        kolbSamplingLUT=0), port, n=100_000)
    _check_guarded(dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.8,
                        useImage=1, exposureControl=0.3), port, n=100_000, image=hex_bokeh_image(255))


def test_thin_retry_kernel_on_random_images(port):
    """The thin-lens retry kernel against the oracle on photograph-like aperture images: coarse value levels (ties, long
    flat stretches of the CDFs -> search brackets far longer than the two entries the fast path handles), black rows and
    columns, ragged sizes on both sides of the 255-column limit of the byte-wide tables, and lens samples at and beyond
    the ends of [0, 1)."""
    from zoic_b200 import ZoicCamera, MODE_GUARDED
    rng = np.random.default_rng(77)
    for w, h in [(255, 255), (200, 77), (33, 140), (7, 5), (255, 511), (256, 64), (300, 40)]:
        levels = int(rng.choice([2, 5, 64]))
        img = (rng.integers(0, levels, (h, w)).astype(np.float32) / np.float32(levels - 1)) ** 2
        img[rng.integers(0, h, max(1, h // 8)), :] = 0.0          # black rows
        img[:, rng.integers(0, w, max(1, w // 8))] = 0.0          # black columns
        if img.sum() == 0:
            img[h // 2, w // 2] = 1.0
        image = np.ascontiguousarray(np.repeat(img[:, :, None], 3, axis=2))
        kw = dict(lensModel=0, focalLength=3.5, fStop=2.0, useImage=1, opticalVignettingDistance=float(rng.uniform(1.0, 4.0)),
                  opticalVignettingRadius=float(rng.uniform(0.7, 1.5)))
        s = random_samples(40_000, seed=w * 1000 + h)
        s[0, 2:] = [0.0, 0.0]
        s[1, 2:] = [1.0, 1.0]
        s[2, 2:] = [0.99999994, 0.5]
        s[3, 2:] = [1.5, 0.25]
        s[4, 2:] = [0.25, -0.25]
        cam = ZoicCamera(image=image, **kw)
        assert cam.mode == MODE_GUARDED
        ref = port.PortCamera(image=image, **kw)
        o, d, st = _run_gpu(cam, s, seed=3, first_index=11)
        o2, d2, st2 = ref.generate(s, seed=3, first_index=11, nthreads=8)
        assert bits_equal(o, o2) and bits_equal(d, d2), (w, h, levels)
        assert st["attempts"] == st2["attempts"] and st["vignetted"] == st2["vignetted"], (w, h)
        cam.close()
        ref.close()


def test_merged_normalisation_equals_the_ieee_operations_for_every_float():
    """The thin-lens retry kernel normalises with one range check around the fast paths of the IEEE root and the IEEE
    reciprocal (lens_math.cuh: normalize_factor).  The device compares it with __frcp_rn(__fsqrt_rn(x)) -- the reference's
    1 / sqrtf(x) in AiV3Normalize -- for all 2^32 bit patterns."""
    import ctypes
    from zoic_b200 import capi
    bad, first = ctypes.c_uint64(123), ctypes.c_uint32(0)
    capi.check(capi.load().zoicb_debug_check_normalize_factor(0, ctypes.byref(bad), ctypes.byref(first)))
    assert bad.value == 0, "first differing bit pattern: 0x%08x" % first.value
    assert first.value == 0xFFFFFFFF


def test_guarded_thin_lens_is_bit_exact(port):
    """The thin lens has no double-precision step, so its default (persistent-warp) kernel keeps the exact
    arithmetic: bit-identical to the oracle, only the order of work differs."""
    from zoic_b200 import ZoicCamera, MODE_GUARDED
    from zoic_b200.synth import hex_bokeh_image
    for image, kw in [(None, dict(lensModel=0, focalLength=3.5, fStop=2.8, opticalVignettingDistance=2.0)),
                      (hex_bokeh_image(255), dict(lensModel=0, focalLength=3.5, fStop=2.8, useImage=1,
                                                  opticalVignettingDistance=2.0, exposureControl=-0.5)),
                      # 255 columns and fewer get byte-wide column tables (camera_state.h: BokehCompact), wider images the
                      # 16-bit ones; 255 x 120: rows and columns differ, 301: the wide path
                      (hex_bokeh_image(255)[:120].copy(), dict(lensModel=0, focalLength=3.5, fStop=2.0, useImage=1,
                                                               opticalVignettingDistance=3.0)),
                      (hex_bokeh_image(301), dict(lensModel=0, focalLength=3.5, fStop=2.8, useImage=1,
                                                  opticalVignettingDistance=2.0)),
                      (None, dict(lensModel=0, focalLength=2.0, fStop=1.4, opticalVignettingDistance=4.0,
                                  opticalVignettingRadius=0.6))]:   # harsh vignetting: many zero-weight rays
        cam = ZoicCamera(image=image, **kw)
        assert cam.mode == MODE_GUARDED
        ref = port.PortCamera(image=image, **kw)
        s = random_samples(300_000, seed=31)
        o, d, st = _run_gpu(cam, s, seed=8, first_index=5)
        o2, d2, st2 = ref.generate(s, seed=8, first_index=5, nthreads=8)
        assert bits_equal(o, o2) and bits_equal(d, d2)
        assert st["attempts"] == st2["attempts"] and st["vignetted"] == st2["vignetted"]
        cam.close()
        ref.close()


# ---------------------------------------------------------------------------------------------------
# committed golden vectors (outputs of the compiled, unmodified reference) and the Arnold-shaped plugin
# ---------------------------------------------------------------------------------------------------
def test_gpu_reproduces_reference_golden_vectors():
    from zoic_b200 import ZoicCamera, MODE_EXACT, MODE_GUARDED
    from zutil import golden_case, golden_names
    for name in golden_names():
        kw, image, meta, s, o_ref, d_ref = golden_case(name)
        cam = ZoicCamera(image=image, mode=MODE_EXACT, **kw)
        o, d, st = _run_gpu(cam, s, seed=meta["seed"], first_index=meta["first_index"])
        assert bits_equal(o, o_ref) and bits_equal(d, d_ref), name
        assert st["attempts"] == meta["stats"]["attempts"], name
        # the default mode on a batch big enough to take the persistent kernels: the golden samples tiled, each
        # copy at its own index (so retries differ), first copy at the golden index
        cam.set_mode(MODE_GUARDED)
        reps = 24
        big = np.tile(s, (reps, 1))
        o, d, _ = _run_gpu(cam, big, seed=meta["seed"], first_index=meta["first_index"])
        res = compare_rays(o[:len(s)], d[:len(s)], o_ref, d_ref)
        assert res["path_flips"] == 0 and res["out_of_tol"] == 0, (name, res)
        cam.close()


def test_arnold_plugin_behind_a_host(port):
    """libzoic_arnold.so driven exactly like the reference plugin: NodeLoader -> Initialize -> Update ->
    CreateRay per sample -> Finish, by the test host that also drives the compiled reference."""
    from oracle import ref
    from zoic_b200 import build
    from zoic_b200.workloads import lens_path
    h, idx = ref.load(plugin=build.PLUGIN)
    assert h.zref_node_name(idx) == b"zoic" and h.zref_node_type(idx) == 0x0002 and h.zref_output_type(idx) == 0xFF
    kw = dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.0, exposureControl=0.4)
    cam = ref.RefCamera(plugin=build.PLUGIN, **kw)
    assert not cam.aborted
    assert "Image distance" in cam.log
    assert h.zref_reverse_ray(cam.c) == 0
    s = random_samples(3000, seed=2)
    o, d, _ = cam.generate(s)                      # the adapter keys retry streams by arrival number, seed 0
    p = port.PortCamera(**kw)
    o2, d2, _ = p.generate(s, seed=0, first_index=0)
    assert bits_equal(o, o2) and bits_equal(d[:, :3], d2[:, :3])
    # derivative workaround of the reference (:1974-1977) for re-sampled rays
    k = int(np.argmax(d2[:, 3] > 0))
    o1, d1, dv = cam.generate_one(s[k], [1, 2, 3, 4])
    cam.close()
    bad = ref.RefCamera(plugin=build.PLUGIN, lensModel=1, lensDataPath="/nonexistent.dat")
    assert bad.aborted and "cannot open lens file" in bad.log
    o, d, _ = bad.generate(s[:4])
    assert (o[:, 3] == 0).all()                    # no camera: zero-weight rays, no crash
    bad.close()
    p.close()


@pytest.mark.parametrize("size", [7, 32, 33, 65, 100])
def test_guarded_kolb_bokeh_image_sizes(port, size):
    """Image sizes that are not 2^k - 1 give the table searches lane-dependent lengths (regression: the
    searches run a warp-uniform number of rounds)."""
    from zoic_b200.synth import hex_bokeh_image
    from zoic_b200.workloads import lens_path
    img = hex_bokeh_image(size)
    if size == 100:
        img = img[:, :77].copy()
    _check_guarded(dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.0,
                        useImage=1), port, n=60_000, image=img)
    _check_exact(dict(lensModel=0, focalLength=3.5, fStop=2.8, opticalVignettingDistance=2.0, useImage=1), port,
                 n=60_000, image=img)


# ---------------------------------------------------------------------------------------------------
# edge cases the reference leaves undefined or accidental
# ---------------------------------------------------------------------------------------------------
def test_edge_samples_follow_the_rulings(port):
    from zoic_b200 import ZoicCamera, MODE_EXACT, MODE_GUARDED
    from zoic_b200.workloads import lens_path
    kw = dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.0)
    s = random_samples(20_000, seed=77)
    s[0] = [0.0, 0.0, 0.3, 0.7]            # film centre: LUT entry 0, no interpolation (the reference steps before begin())
    s[1] = [0.25, -0.1, 0.5, 0.5]          # lens square centre: 0/0 in the concentric map -> NaN ray, weight 1 (kept)
    s[2] = [0.125 / 1.8, 0.0, 0.2, 0.9]    # film radius exactly on a LUT key
    s[3] = [1.0, 2.0 / 3.0, 0.0, 0.0]      # corners of both squares
    s[4] = [-1.0, -2.0 / 3.0, 1.0, 1.0]    # lens sample exactly 1.0 (a retry draw can round up to it)
    ref = port.PortCamera(**kw)
    o2, d2, _ = ref.generate(s, seed=5, first_index=0)
    assert np.isnan(o2[1, :3]).all() and o2[1, 3] == 1.0 and d2[1, 3] == 0.0
    for mode in (MODE_EXACT, MODE_GUARDED):
        cam = ZoicCamera(mode=mode, **kw)
        o, d, _ = _run_gpu(cam, s, seed=5, first_index=0)
        assert np.isnan(o[1, :3]).all() and o[1, 3] == 1.0 and d[1, 3] == 0.0
        keep = np.ones(len(s), bool)
        keep[1] = False
        if mode == MODE_EXACT:
            assert bits_equal(o[keep], o2[keep]) and bits_equal(d[keep], d2[keep])
        else:
            res = compare_rays(o[keep], d[keep], o2[keep], d2[keep])
            assert res["path_flips"] == 0 and res["out_of_tol"] == 0, res
        cam.close()
    ref.close()


def test_film_radius_beyond_the_lut_clamps(port):
    """sensorWidth 9 cm puts film points beyond the last LUT key (3.875 cm): last entry, extrapolated (ruling)."""
    from zoic_b200.workloads import lens_path
    _check_exact(dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.0,
                      sensorWidth=9.0, sensorHeight=6.0), port, n=40_000)
    _check_guarded(dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.0,
                        sensorWidth=9.0, sensorHeight=6.0), port, n=100_000)


def test_large_batch_spans_many_chunks_and_keeps_counters(port):
    """2^22 samples through the persistent kernels: every sample written exactly once, counters add up."""
    from zoic_b200 import ZoicCamera
    from zoic_b200.workloads import config4
    wl = config4()
    cam = ZoicCamera(**wl.params)
    n = 1 << 22
    s = cam.synth_samples(*wl.synth_args(), 0, n)
    rays = torch.full((n, 8), float("nan"), device="cuda")
    cam.reset_stats()
    cam.create_rays(s, seed=wl.seed, first_index=0, out=rays)
    torch.cuda.synchronize()
    o, d = rays[:, :4], rays[:, 4:]
    st = cam.stats()
    assert not torch.isnan(o[:, 3]).any() and not torch.isnan(d[:, 3]).any()
    assert st["rays"] == n and st["success"] + st["vignetted"] == n
    assert st["success"] == int((o[:, 3] != 0).sum()) and st["vignetted"] == int((o[:, 3] == 0).sum())
    assert st["attempts"] == n + int(d[:, 3].sum(dtype=torch.float64))
    cam.close()


def test_random_cameras_sweep(port):
    """Random node parameters over their documented ranges (src/zoic.mtd): EXACT is bit-identical, GUARDED has
    zero path flips and stays inside the tolerance, for every lens table and both lens models."""
    from zoic_b200 import ZoicCamera, MODE_EXACT, MODE_GUARDED
    from zoic_b200.workloads import LENSES, lens_path
    rng = np.random.default_rng(2024)
    lenses = sorted(LENSES)
    worst = {"o": 0.0, "d": 0.0}
    for trial in range(24):
        if trial % 4 == 3:
            kw = dict(lensModel=0, focalLength=float(rng.uniform(1.5, 10.0)), fStop=float(rng.uniform(1.0, 16.0)),
                      focalDistance=float(rng.uniform(20.0, 500.0)),
                      opticalVignettingDistance=float(rng.choice([0.0, rng.uniform(0.5, 6.0)])),
                      opticalVignettingRadius=float(rng.uniform(0.5, 2.0)), exposureControl=float(rng.uniform(-2, 2)))
        else:
            lens = lenses[trial % len(lenses)]
            native = 1.0 if "fisheye" in lens else 5.0
            kw = dict(lensModel=1, lensDataPath=lens_path(lens), focalLength=float(native * rng.uniform(0.6, 1.8)),
                      fStop=float(rng.uniform(1.2, 11.0)), focalDistance=float(rng.uniform(25.0, 800.0)),
                      kolbSamplingLUT=int(rng.random() < 0.8), exposureControl=float(rng.uniform(-1, 1)),
                      sensorWidth=float(rng.choice([3.6, 2.4, 1.8])))
        ref = port.PortCamera(**kw)
        s = random_samples(60_000, seed=100 + trial)
        o2, d2, st2 = ref.generate(s, seed=trial, first_index=1 << 33, nthreads=8)
        for mode in (MODE_EXACT, MODE_GUARDED):
            cam = ZoicCamera(mode=mode, **kw)
            o, d, st = _run_gpu(cam, s, seed=trial, first_index=1 << 33)
            if mode == MODE_EXACT:
                assert bits_equal(o, o2) and bits_equal(d, d2), kw
            else:
                res = compare_rays(o, d, o2, d2)
                assert res["path_flips"] == 0 and res["out_of_tol"] == 0, (kw, res)
                worst["o"] = max(worst["o"], res["max_origin_err"])
                worst["d"] = max(worst["d"], res["max_dir_err"])
            assert st["attempts"] == st2["attempts"] and st["vignetted"] == st2["vignetted"], kw
            cam.close()
        ref.close()
    assert worst["o"] < 1e-5 and worst["d"] < 1e-5


def test_camera_to_world_epilogue_matches_the_contract(port):
    """zoicb_transform_rays against the CPU statement of its arithmetic, bit for bit; in place and
    out of place; weight / tries untouched; identity matrix is the identity; empty batch is a no-op."""
    from zoic_b200 import ZoicCamera
    from zoic_b200.workloads import config2
    wl = config2()
    cam = ZoicCamera(**wl.params)
    n = 300_001   # not a multiple of anything
    s = cam.synth_samples(*wl.synth_args(), 5_000_000, n)
    rays = cam.create_rays(s, seed=wl.seed, first_index=5_000_000)
    torch.cuda.synchronize()
    host = rays.cpu().numpy()
    rng = np.random.default_rng(3)
    # a rigid camera placement (rotation about a skew axis + translation in cm) and a general affine matrix
    a = rng.normal(size=3); a /= np.linalg.norm(a); th = 0.7
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    rigid = np.concatenate([R, np.array([[120.0], [-35.5], [900.25]])], axis=1).astype(np.float32)
    general = rng.normal(size=(3, 4)).astype(np.float32)
    for m in (rigid, general):
        want = port.transform_rays(host, m)
        got = cam.transform_rays(rays, m)
        torch.cuda.synchronize()
        assert bits_equal(got.cpu().numpy(), want)
        assert bits_equal(got.cpu().numpy()[:, 3], host[:, 3]) and bits_equal(got.cpu().numpy()[:, 7], host[:, 7])
    # rigid motion keeps directions unit length (to the fp32 rounding of the matrix and of the directions themselves)
    g = cam.transform_rays(rays, rigid).cpu().numpy()
    live = host[:, 3] != 0
    assert np.abs(np.linalg.norm(g[live, 4:7].astype(np.float64), axis=1) - 1.0).max() < 5e-6
    ident = np.eye(3, 4, dtype=np.float32)
    same = cam.transform_rays(rays, ident).cpu().numpy()
    # identity: fma(1, x, fma(0, y, fma(0, z, 0))) = x exactly (also for -0.0 + 0.0 = +0.0: compare values)
    assert np.array_equal(same[:, :3], host[:, :3]) and np.array_equal(same[:, 4:7], host[:, 4:7])
    inplace = rays.clone()
    cam.transform_rays(inplace, general, out=inplace)
    torch.cuda.synchronize()
    assert bits_equal(inplace.cpu().numpy(), port.transform_rays(host, general))
    cam.transform_rays(rays[:0], general)
    cam.close()


@pytest.mark.parametrize("pool", ["2", "3"])
def test_both_pool_flavours_on_every_lens(pool):
    """The host picks the plain or the rim pre-test flavour of the pool kernel per camera; ZOICB_POOL=2 / 3 forces one of
    them (read once per process), so the all-lenses and bokeh parity tests run again in a child process for each."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, ZOICB_POOL=pool)
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity.py"), "-q", "-x", "-m", "gpu", "-k",
                        "guarded_kolb_all_lenses or guarded_kolb_no_lut_and_bokeh or guarded_kolb_bokeh_image_sizes or "
                        "large_batch_spans_many_chunks"], env=env, capture_output=True, text=True, cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("case", ["dg_f28_focus23", "dg_nolut", "fisheye"])
def test_draw_file_matches_the_reference_draw_build(case, tmp_path):
    """zoicb_write_draw_file against draw.zoic files written by the -D_DRAW build of the unmodified
    reference (tests/golden/draw_rays_*.txt, tools/make_golden_draw.py) for the samples that build drew: byte for byte,
    header and every attempt's (z, y) path."""
    import os
    from zoic_b200 import ZoicCamera
    from zoic_b200.workloads import lens_path
    sys_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_draw", os.path.join(sys_path, "make_golden_draw.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    kw = dict(mg.CASES[case])
    kw["lensDataPath"] = lens_path(kw["lensDataPath"])
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    data = np.load(os.path.join(gold, "draw_rays_%s.npz" % case))
    want = open(os.path.join(gold, "draw_rays_%s.txt" % case)).read()
    cam = ZoicCamera(lensModel=1, **kw)
    out = tmp_path / "draw.zoic"
    cam.write_draw_file(str(out), data["samples"], seed=int(data["seed"]), indices=data["index"])
    got = open(out).read()
    cam.close()
    assert got.split("RAYS{")[0] == want.split("RAYS{")[0]          # header
    assert got == want


# ---------------------------------------------------------------------------------------------------
# the image-based aperture tables, built on the GPU (bokeh_build.cu)
# ---------------------------------------------------------------------------------------------------
def _bokeh_images():
    from zoic_b200.synth import hex_bokeh_image
    rng = np.random.default_rng(17)
    photo = np.round(rng.random((96, 160, 3)) * 15).astype(np.float32) / 15   # 16 grey levels: ties everywhere
    photo[rng.random((96, 160)) < 0.3] = 0.0
    return {
        "hex255": hex_bokeh_image(255), "hex33-cropped": hex_bokeh_image(33)[:, :20].copy(), "hex32": hex_bokeh_image(32),
        "flat": np.ones((7, 9, 3), np.float32), "one-row": rng.random((1, 16, 4)).astype(np.float32),
        "one-column": rng.random((16, 1, 3)).astype(np.float32), "black": np.zeros((5, 5, 3), np.float32),
        "quantised-photo": photo, "wide": rng.random((3, 5000, 3)).astype(np.float32),
        "tall": rng.random((700, 40, 3)).astype(np.float32),
    }


def _same_table(a, b):
    """Bit equality; NaNs (a black image: 0 * inf) only have to be NaNs -- x86 and the GPU sign them differently."""
    if a.dtype != b.dtype or a.shape != b.shape:
        return False
    if a.dtype == np.float32:
        nan = np.isnan(a)
        return np.array_equal(nan, np.isnan(b)) and np.array_equal(a[~nan].view(np.uint32), b[~nan].view(np.uint32))
    return np.array_equal(a, b)


@pytest.mark.parametrize("name", ["hex255", "hex33-cropped", "hex32", "flat", "one-row", "one-column", "black",
                                  "quantised-photo", "wide", "tall"])
def test_bokeh_tables_built_on_the_gpu_equal_the_oracle_tables(port, name):
    """cdfRow, rowIndices, cdfColumn, columnIndices entry for entry (float bits, tie order of the sorts included)
    against the oracle; and a camera created with the image carries the same tables."""
    from zoic_b200 import ZoicCamera, build_bokeh_tables
    img = _bokeh_images()[name]
    kw = dict(lensModel=0, focalLength=3.5, fStop=2.8, useImage=1)
    p = port.PortCamera(image=img, **kw)
    want = p.bokeh_tables()
    p.close()
    got, ms = build_bokeh_tables(img)
    assert ms > 0
    for a, b in zip(got, want):
        assert _same_table(a, b)
    cam = ZoicCamera(image=img, **kw)
    for a, b in zip(cam.bokeh_tables(), want):
        assert _same_table(a, b)
    cam.close()


def test_bokeh_table_build_large_image_against_the_host_statement():
    """1024 x 768 photograph-like image (256 grey levels, so thousands of ties per row) against the host statement of
    the build, which the CPU suite pins to the oracle."""
    from zoic_b200 import build_bokeh_tables, host_setup
    rng = np.random.default_rng(23)
    img = (rng.integers(0, 256, (768, 1024, 3)) / 255.0).astype(np.float32)
    _, want = host_setup(image=img, lensModel=0, focalLength=3.5, fStop=2.8, useImage=1)
    got, ms = build_bokeh_tables(img)
    for a, b in zip(got, want):
        assert _same_table(a, b)


# ---------------------------------------------------------------------------------------------------
# BASELINE.json's full sizes: size-independent properties + oracle windows inside the big batch
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["headline", "config3", "config4"])
def test_full_size_batches_hold_the_invariants(port, name):
    """One zoicb_generate over the whole configuration (2.12 G samples; config 4 is halved until it fits the device).
    Checked: every record written once, counters add up, weights are 0 or the exposure scale, live directions are unit
    vectors, tries <= 26; three 2^16-sample windows (start, middle, end) equal their own small launches bit for bit
    (batch boundaries do not matter) and match the oracle (bit-exact for the thin lens, zero path flips and the 1e-5
    tolerance for the guarded raytraced lens)."""
    from zoic_b200 import ZoicCamera
    from zoic_b200 import workloads
    wl = workloads.BY_NAME[name]()
    n = wl.n
    torch.cuda.empty_cache()   # blocks cached by earlier tests count as used in mem_get_info
    free, _ = torch.cuda.mem_get_info()
    while n * 48 > free * 0.85:
        n //= 2
    if n < (1 << 28):
        pytest.skip("not enough device memory for a full-size batch")
    image = wl.image()
    cam = ZoicCamera(image=image, **wl.params)
    s = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    tile = 1 << 27
    for b in range(0, n, tile):
        m = min(tile, n - b)
        cam.synth_samples(*wl.synth_args(), b, m, out=s[b:b + m])
    rays = torch.full((n, 8), float("nan"), device="cuda")
    cam.reset_stats()
    cam.create_rays(s, seed=wl.seed, first_index=0, out=rays)
    torch.cuda.synchronize()
    st = cam.stats()
    live = tries_sum = 0
    worst = 0.0
    for b in range(0, n, tile):
        r = rays[b:b + tile]
        w, t = r[:, 3], r[:, 7]
        assert not torch.isnan(w).any() and not torch.isnan(t).any()
        assert bool(((w == 0) | (w == 1)).all()) and float(t.max()) <= 26 and float(t.min()) >= 0
        alive = w != 0
        live += int(alive.sum())
        tries_sum += int(t.sum(dtype=torch.float64))
        d = r[:, 4:7]
        err = ((d * d).sum(1) - 1).abs()
        worst = max(worst, float(err[alive].max()) if bool(alive.any()) else 0.0)
        assert bool(torch.isfinite(r[:, :7][alive]).all())
    assert st["rays"] == n and st["success"] + st["vignetted"] == n
    assert st["success"] == live and st["attempts"] == n + tries_sum
    assert worst <= 1e-5, worst   # |d|^2 - 1 of live rays
    ref = port.PortCamera(image=image, **wl.params)
    k = 1 << 16
    for a in (0, (n // 2) - 12345, n - k):
        win = s[a:a + k].contiguous()
        again = cam.create_rays(win, seed=wl.seed, first_index=a)
        torch.cuda.synchronize()
        assert torch.equal(again.view(torch.int32), rays[a:a + k].view(torch.int32))
        g = again.cpu().numpy()
        o_ref, d_ref, _ = ref.generate(win.cpu().numpy(), seed=wl.seed, first_index=a, nthreads=8)
        res = compare_rays(g[:, :4], g[:, 4:], o_ref, d_ref)
        assert res["path_flips"] == 0 and res["out_of_tol"] == 0, res
        if wl.params["lensModel"] == 0:
            assert bits_equal(g[:, :4], o_ref) and bits_equal(g[:, 4:], d_ref)
    cam.close()
    ref.close()
    del s, rays
    torch.cuda.empty_cache()
