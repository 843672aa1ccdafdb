"""GPU tests of the whole-frame job runner (zoicb_run_job), the on-device parity census, the device-side exit-pupil LUT
boxes and the NVLink gather, all through the C ABI.

What is compared with what:
  * windows of records copied out of a streamed job  vs  the CPU oracle on the same samples (bit-exact for the thin lens;
    zero path flips and the 1e-5 north-star tolerance for the guarded raytraced lens) and vs a direct zoicb_generate of
    the same samples (bit for bit: batch and tile boundaries do not matter);
  * the job's checksum / counts  vs  numpy over a direct zoicb_generate of the whole frame;
  * the census (GUARDED vs EXACT, every record, on the device)  vs  zero flips, nothing out of tolerance;
  * BASELINE.json's full sizes (headline 2.1 G, config 4 4.2 G, config 5 34 G samples per lens) streamed at full size.
"""
import os

import numpy as np
import pytest

from zutil import bits_equal, compare_rays

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

K = np.array([0x9E3779B1, 0x85EBCA77, 0xC2B2AE3D, 0x27D4EB2F, 0x165667B1, 0xD3A2646D, 0xFD7046C5, 0xB55A4F09], np.uint64)


def numpy_checksum(rays):
    """zoic_b200/csrc/job.cu consume_rays_kernel: sum over records of sum_j word_j * K_j mod 2^64 (+ zero weights, tries)."""
    w = np.ascontiguousarray(rays, np.float32).view(np.uint32).astype(np.uint64)
    with np.errstate(over="ignore"):
        s = int((w * K[None, :]).sum(dtype=np.uint64))
    return s, int((rays[:, 3] == 0).sum()), int(rays[:, 7].astype(np.float64).sum())


def small_frame(kind):
    from zoic_b200.synth import hex_bokeh_image
    from zoic_b200.workloads import Workload, _kolb
    if kind == "kolb":
        return Workload("small double gauss", 256, 144, 64, 41, _kolb("double_gauss_f2.0.dat", 5.0, 2.0)), None
    if kind == "fisheye":
        return Workload("small fisheye", 256, 144, 64, 42, _kolb("fisheye_muller_f4.0.dat", 1.0, 4.0)), None
    wl = Workload("small thin + OV + hex", 256, 144, 64, 43, dict(lensModel=0, focalLength=3.5, fStop=2.8, useImage=1,
                                                                   opticalVignettingDistance=2.0, opticalVignettingRadius=1.0))
    return wl, hex_bokeh_image(255)


def check_windows(port, wl, image, cam, res, windows, count):
    """the job's windows against a direct generate of the same samples (bits) and against the oracle"""
    ref = port.PortCamera(image=image, **wl.params)
    got = res["windows"].cpu().numpy()
    for k, a in enumerate(windows):
        s = cam.synth_samples(*wl.synth_args(), a, count)
        direct = cam.create_rays(s, seed=wl.seed, first_index=a)
        torch.cuda.synchronize()
        assert bits_equal(got[k], direct.cpu().numpy()), (wl.name, a)
        o_ref, d_ref, _ = ref.generate(s.cpu().numpy(), seed=wl.seed, first_index=a, nthreads=8)
        r = compare_rays(got[k][:, :4], got[k][:, 4:], o_ref, d_ref)
        assert r["path_flips"] == 0 and r["out_of_tol"] == 0, (wl.name, a, r)
        if wl.params["lensModel"] == 0:
            assert bits_equal(got[k][:, :4], o_ref) and bits_equal(got[k][:, 4:], d_ref)
    ref.close()


@pytest.mark.parametrize("kind", ["kolb", "fisheye", "thin"])
def test_streamed_job_equals_one_big_generate(port, kind):
    from zoic_b200 import ZoicCamera
    wl, image = small_frame(kind)
    cam = ZoicCamera(image=image, **wl.params)
    n, tile, wc = wl.n, 1 << 18, 1 << 14
    windows = [0, 3 * tile - wc // 2, n - wc]        # start, across a tile boundary, end
    res = cam.run_job(*wl.synth_args(), wl.seed, 0, n, tile=tile, census=(kind != "thin"), windows=windows, window_count=wc)
    assert res["rays"] == n and res["tiles"] == (n + tile - 1) // tile and res["consumed"] == n
    s = cam.synth_samples(*wl.synth_args(), 0, n)
    cam.reset_stats()
    whole = cam.create_rays(s, seed=wl.seed, first_index=0)
    torch.cuda.synchronize()
    st = cam.stats()
    want = numpy_checksum(whole.cpu().numpy())
    assert (res["checksum"], res["zero_weight"], res["tries_sum"]) == want
    for k in ("rays", "success", "vignetted", "attempts", "element_visits", "total_internal_reflection"):
        assert res["stats"][k] == st[k], k
    assert res["stats"]["vignetted"] == res["zero_weight"] and res["stats"]["attempts"] == n + res["tries_sum"]
    if kind != "thin":
        assert res["census_rays"] == n and res["census_flips"] == 0 and res["census_out_of_tol"] == 0
        assert res["census_live"] == st["success"] and res["census_max_rel_origin"] <= 1e-5 and res["census_max_dir"] <= 1e-5
        assert res["census_stats"]["attempts"] == st["attempts"]          # the EXACT pass walked the same paths
    check_windows(port, wl, image, cam, res, windows, wc)
    # one stream, no overlap: the same records
    ser = cam.run_job(*wl.synth_args(), wl.seed, 0, n, tile=tile, serial=True)
    assert ser["checksum"] == res["checksum"] and ser["consumed"] == n
    # a share in the middle of the frame (what a rank of a multi-GPU job runs), ragged last tile
    first, count = n // 4 + 12345, n // 2 + 777
    part = cam.run_job(*wl.synth_args(), wl.seed, first, count, tile=tile)
    assert part["checksum"] == numpy_checksum(whole[first:first + count].cpu().numpy())[0] and part["rays"] == count
    cam.close()


def test_census_counts_what_it_should(port):
    """zoicb_census on doctored buffers: a changed weight / tries is a flip, a moved live ray is out of tolerance, a moved
    zero-weight record is neither; NaN rays with equal NaN pattern are equal."""
    from zoic_b200 import ZoicCamera, MODE_EXACT
    wl, _ = small_frame("kolb")
    cam = ZoicCamera(**wl.params)
    n = 200_000
    s = cam.synth_samples(*wl.synth_args(), 0, n)
    fast = cam.create_rays(s, seed=wl.seed)
    cam.set_mode(MODE_EXACT)
    exact = cam.create_rays(s, seed=wl.seed)
    torch.cuda.synchronize()
    base = cam.census(fast, exact)
    assert base["rays"] == n and base["flips"] == 0 and base["out_of_tol"] == 0 and 0 < base["max_dir"] <= 1e-5
    live = torch.nonzero(exact[:, 3] != 0)[:, 0]
    dead = torch.nonzero(exact[:, 3] == 0)[:, 0]
    assert len(live) > 100 and len(dead) > 10
    bad = fast.clone()
    bad[live[0], 3] = 0.0                      # weight flip
    bad[live[1], 7] += 1.0                     # tries flip
    bad[live[2], 0] += 1e-3                    # origin off by 1e-3 cm
    bad[live[3], 5] += 1e-4                    # direction off
    bad[dead[0], 0] += 5.0                     # zero-weight records carry no ray
    bad[live[4], :3] = float("nan")            # NaN on one side only
    r = cam.census(bad, exact)
    assert r["flips"] == 2 and r["out_of_tol"] == 3 and r["max_rel_origin"] > 1e-5
    both = exact.clone()
    both[live[5], :3] = float("nan")
    bad2 = fast.clone()
    bad2[live[5], :3] = float("nan")
    r = cam.census(bad2, both)
    assert r["flips"] == 0 and r["out_of_tol"] == 0
    cam.close()


# ---------------------------------------------------------------------------------------------------
# ray differentials: the derivatives the reference leaves as a TODO (src/zoic.cpp:12-13, :1971-1977)
# ---------------------------------------------------------------------------------------------------
def test_ray_differentials_match_the_contract(port):
    """zoicb_differentials against the CPU statement of its contract (oracle/zoic_port.cpp: zport_differentials), bit for
    bit, for both lens models, with and without LUT / image / depth of field; rays from the default (GUARDED) mode and
    from EXACT mode give the same differentials (they agree on weight and tries); zero-weight rays get zeros; a zero step
    gives zeros; a thin lens without depth of field has the analytic pinhole derivative."""
    from zoic_b200 import ZoicCamera, MODE_EXACT
    from zoic_b200.synth import hex_bokeh_image
    from zoic_b200.workloads import lens_path
    cases = [
        (dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.0), None),
        (dict(lensModel=1, lensDataPath=lens_path("fisheye_muller_f4.0.dat"), focalLength=1.0, fStop=4.0), None),
        (dict(lensModel=1, lensDataPath=lens_path("tessar_f2.8.dat"), focalLength=5.0, fStop=2.8, kolbSamplingLUT=0, useImage=1), 65),
        (dict(lensModel=0, focalLength=3.5, fStop=2.8, opticalVignettingDistance=2.0, useImage=1), 255),
        (dict(lensModel=0, focalLength=3.5, fStop=2.8, opticalVignettingDistance=4.0, opticalVignettingRadius=0.6), None),
        (dict(lensModel=0, focalLength=3.5, fStop=2.8, useDof=0), None),
    ]
    W, H, spp, n, first = 640, 360, 16, 150_001, 77_777
    dsx, dsy = 2.0 / W, 2.0 / W            # one pixel in screen space (sy spans the aspect-scaled range at the same pitch)
    for kw, img in cases:
        image = hex_bokeh_image(img) if img else None
        cam = ZoicCamera(image=image, **kw)
        ref = port.PortCamera(image=image, **kw)
        s = cam.synth_samples(W, H, spp, 9, first, n)
        rays = cam.create_rays(s, seed=5, first_index=first)
        got = cam.differentials(s, rays, dsx, dsy, seed=5, first_index=first)
        torch.cuda.synchronize()
        hs, hr = s.cpu().numpy(), rays.cpu().numpy()
        want = ref.differentials(hs, hr, dsx, dsy, seed=5, first_index=first)
        assert bits_equal(got.cpu().numpy(), want), kw
        dead = hr[:, 3] == 0
        assert not want[dead].any()
        live = ~dead
        assert np.isfinite(want).all() and np.abs(want[live, 6:]).max() < 0.05     # a pixel's worth of direction change
        assert (np.abs(want[live, 6:9]).sum(1) > 0).mean() > 0.95                   # and almost never stopped
        # camera -> world: every differential vector times the 3x3 part of the matrix, bit for bit; in place too
        m = np.random.default_rng(7).normal(size=(3, 4)).astype(np.float32)
        world = cam.transform_differentials(got, m)
        assert bits_equal(world.cpu().numpy(), port.transform_differentials(want, m))
        same = got.clone()
        cam.transform_differentials(same, m, out=same)
        assert torch.equal(same.view(torch.int32), world.view(torch.int32))
        cam.set_mode(MODE_EXACT)
        rays_x = cam.create_rays(s, seed=5, first_index=first)
        got_x = cam.differentials(s, rays_x, dsx, dsy, seed=5, first_index=first)
        assert torch.equal(got.view(torch.int32), got_x.view(torch.int32))
        assert not cam.differentials(s, rays, 0.0, 0.0, seed=5, first_index=first).any()
        if kw.get("useDof", 1) == 0:
            # pinhole: dir = normalize(sx t, sy t, 1) flipped in z; d dir / d sx by forward difference in float64
            c = cam.constants()
            t = float(c["tan_fov"])
            p0 = np.stack([hs[:, 0].astype(np.float64) * t, hs[:, 1].astype(np.float64) * t, np.ones(n)], 1)
            p1 = p0.copy()
            p1[:, 0] = (hs[:, 0].astype(np.float64) + dsx) * t
            d0 = p0 / np.linalg.norm(p0, axis=1, keepdims=True)
            d1 = p1 / np.linalg.norm(p1, axis=1, keepdims=True)
            an = (d1 - d0) * np.array([1, 1, -1.0])
            assert np.abs(want[:, 6:9] - an).max() < 5e-7 and not want[:, :6].any()
        cam.close()
        ref.close()


# ---------------------------------------------------------------------------------------------------
# BASELINE.json's full sizes, streamed ("Config 4/5 ... streamed in tiles")
# ---------------------------------------------------------------------------------------------------
def _full_size_names():
    from zoic_b200 import workloads
    return ["headline", "config4"] + workloads.CONFIG5


@pytest.mark.parametrize("name", _full_size_names())
def test_full_size_streamed_jobs(port, name):
    """Every sample of the configuration through zoicb_run_job (tiles of 2^28 synthesised on the device, generated,
    consumed): all records consumed, counters consistent; three 2^16-sample windows (start, a tile boundary in the middle,
    end) equal to their own small launches bit for bit and to the oracle; the census of the first 2.1 G rays (GUARDED vs
    EXACT, every record): no flips, nothing out of tolerance."""
    from zoic_b200 import ZoicCamera, workloads
    wl = workloads.BY_NAME[name]()
    n = wl.n
    if int(os.environ.get("ZOICB_TEST_MAX_SAMPLES", "0")):
        n = min(n, int(os.environ["ZOICB_TEST_MAX_SAMPLES"]))
    cam = ZoicCamera(**wl.params)
    tile, wc = 1 << 28, 1 << 16
    mid = (n // tile // 2) * tile
    windows = [0, max(mid, wc) - wc // 2, n - wc]
    res = cam.run_job(*wl.synth_args(), wl.seed, 0, n, tile=tile, windows=windows, window_count=wc)
    st = res["stats"]
    assert res["rays"] == n and res["consumed"] == n and st["rays"] == n and st["success"] + st["vignetted"] == n
    assert st["vignetted"] == res["zero_weight"] and st["attempts"] == n + res["tries_sum"]
    assert st["exact_reruns"] < 0.03 * n
    check_windows(port, wl, None, cam, res, windows, wc)
    m = min(n, 2_123_366_400)
    cen = cam.run_job(*wl.synth_args(), wl.seed, 0, m, tile=tile, census=True)
    assert cen["census_rays"] == m and cen["census_flips"] == 0 and cen["census_out_of_tol"] == 0, cen
    assert cen["census_max_rel_origin"] <= 1e-5 and cen["census_max_dir"] <= 1e-5
    assert cen["census_stats"]["attempts"] == cen["stats"]["attempts"] and cen["census_stats"]["vignetted"] == cen["stats"]["vignetted"]
    print("\n%s: %d rays in %.1f ms (%.0f Mrays/s), census of %d rays: 0 flips, max rel origin %.2e, max dir %.2e, "
          "%.2f %% exact re-runs" % (name, n, res["device_ms"], n / res["device_ms"] / 1e3, m, cen["census_max_rel_origin"],
                                     cen["census_max_dir"], 100.0 * st["exact_reruns"] / n))
    cam.close()


# ---------------------------------------------------------------------------------------------------
# exit-pupil LUT boxes on the device (reference src/zoic.cpp:1391-1452)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lens", ["double_gauss_f2.0.dat", "fisheye_muller_f4.0.dat", "petzval_f1.6.dat", "telephoto_f5.0.dat",
                                  "tessar_f2.8.dat", "triplet_f2.5.dat", "mori_f2.8.dat", "petzval_f1.25.dat"])
def test_lut_boxes_built_on_the_gpu_equal_the_oracle(port, lens):
    """lutMinX/Y, lutMaxX/Y of all 32 film positions -- 3.2 M candidates classified and folded on the GPU -- bit for bit
    against the oracle's node_update; and what creation costs (the reference spends 0.4-0.7 s here)."""
    from zoic_b200 import ZoicCamera
    from zoic_b200.workloads import LENSES, lens_path
    fnum, focal = LENSES[lens]
    kw = dict(lensModel=1, lensDataPath=lens_path(lens), focalLength=focal, fStop=fnum)
    cam = ZoicCamera(**kw)
    ref = port.PortCamera(**kw)
    got, want = cam.constants()["lut"], ref.constants()["lut"]
    assert got.shape == (32, 5) and bits_equal(got, want)
    t = cam.create_times()
    assert 0 < t["lut_ms"] <= t["create_ms"]
    print("\n%s: zoicb_create %.1f ms (exit-pupil LUT %.1f ms)" % (lens, t["create_ms"], t["lut_ms"]))
    cam.close()
    ref.close()


def test_lut_box_fold_rearm_quirk_on_the_gpu():
    """The reference re-arms a film position's box whenever min.x + min.y is exactly 0 (src/zoic.cpp:1423) -- always at
    the first accepted candidate, and again if the minima cancel later.  Candidates crafted to cancel (u = 1/4 and 3/4
    give -ap/2 and +ap/2) at many places, also next to chunk boundaries of the device kernel: device fold == host fold."""
    from zoic_b200 import debug_lut_boxes
    rng = np.random.default_rng(5)
    n_film, per_film, ap = 12, 5000, 1.7
    draws = rng.integers(0, 1 << 32, size=(n_film, per_film, 2), dtype=np.uint64).astype(np.uint32)
    accept = (rng.random((n_film, per_film)) < 0.3).astype(np.uint8)
    quarter, three = np.uint32(1 << 30), np.uint32(3 << 30)
    for f in range(n_film):
        for pos in rng.integers(0, per_film, size=f * 3):     # film 0: no crafted candidate at all
            draws[f, pos] = (quarter, three) if rng.random() < 0.5 else (three, quarter)
            accept[f, pos] = 1
    for pos in (0, 31, 32, 63, 64, per_film - 1):             # chunk edges of the 32-wide device fold
        draws[5, pos] = (quarter, three)
        accept[5, pos] = 1
    accept[7, :] = 0                                           # nothing accepted: the box stays at the origin
    # films 8..11: a LATE re-arm that changes the final box.  Only candidates with x > -0.2 ap and y > 0.6 ap are accepted,
    # so the crafted candidate (-ap/2, +ap/2) becomes the minimum in x AND in y; the minima cancel and the next accepted
    # candidate re-arms the box, which forgets every extreme seen before (lane positions around the device kernel's chunks)
    u = draws.astype(np.float32) * np.float32(2.0 ** -32)
    p = (u * np.float32(2.0) - np.float32(1.0)) * np.float32(ap)
    for f, pos in zip(range(8, 12), (2500, 2528, 2559, 2560)):
        accept[f] = ((p[f, :, 0] > -0.2 * ap) & (p[f, :, 1] > 0.6 * ap)).astype(np.uint8)
        draws[f, pos] = (quarter, three)
        accept[f, pos] = 1
    dev, host = debug_lut_boxes(draws, accept, n_film, per_film, ap, device=0)
    assert bits_equal(dev, host)
    assert not host[7].any() and host[1:7].any()
    # the late re-arm really happened: the crafted candidate is the plain minimum in x, the fold has forgotten it
    u = draws.astype(np.float32) * np.float32(2.0 ** -32)
    p = (u * np.float32(2.0) - np.float32(1.0)) * np.float32(ap)
    plain_minx = np.where(accept == 1, p[..., 0], np.inf).min(1)
    assert (plain_minx[8:12] == np.float32(-0.5) * np.float32(ap)).all() and (host[8:12, 0] > plain_minx[8:12]).all()


# ---------------------------------------------------------------------------------------------------
# the NVLink gather: two (or more) GPUs, one process each
# ---------------------------------------------------------------------------------------------------
def _gather_worker(rank, world, port_no, transport, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from zoic_b200 import Gather, ZoicCamera
    from zoic_b200.distributed import connect_gather, job_share
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl, _ = small_frame("kolb")
    cam = ZoicCamera(device=rank, **wl.params)
    first, count = job_share(wl.n, wl.passes, rank, world)
    tile = 1 << 17                                   # 9 rounds for two ranks
    rounds = (count + tile - 1) // tile
    g = Gather(rank, rank, world, 0, tile, slots=rounds, transport=transport)   # slots = rounds: rank 0 keeps the whole job
    connect_gather(g)
    res = None
    for _ in range(2):                               # twice: the flags carry on counting across jobs
        dist.barrier()
        res = cam.run_job(*wl.synth_args(), wl.seed, first, count, gather=g, gather_counts=[count] * world)
    dist.barrier()
    if rank == 0:
        recs = []
        for r in range(world):
            for k in range(rounds):
                m = min(tile, count - k * tile)
                recs.append(g.read(k, r, 0, m))
        np.savez(os.path.join(out_dir, "gathered_%s.npz" % transport), rays=np.concatenate(recs),
                 totals=np.array([res["checksum"], res["zero_weight"], res["tries_sum"], res["consumed"]], np.uint64))
    dist.barrier()
    g.close()
    cam.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["fused", "push", "nccl"])
def test_gather_to_consumer_over_nvlink(tmp_path, transport):
    """Two ranks generate their shares of one frame; every record lands in rank 0's round buffers (kernels storing over
    NVLink / copy-engine push / ncclSend-Recv) and equals, bit for bit and in job order, the frame generated on one GPU;
    the consumer's checksum equals numpy's over that frame."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import socket
    import torch.multiprocessing as mp
    from zoic_b200 import ZoicCamera
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    world = 2
    mp.spawn(_gather_worker, args=(world, port_no, transport, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered_%s.npz" % transport))
    wl, _ = small_frame("kolb")
    cam = ZoicCamera(**wl.params)
    smp = cam.synth_samples(*wl.synth_args(), 0, wl.n)
    whole = cam.create_rays(smp, seed=wl.seed, first_index=0)
    torch.cuda.synchronize()
    whole = whole.cpu().numpy()
    assert bits_equal(got["rays"], whole)
    want = numpy_checksum(whole)
    assert tuple(int(x) for x in got["totals"]) == want + (wl.n,)
    cam.close()
