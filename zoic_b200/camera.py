"""Host-side mirror of the reference camera node on top of the C ABI.

`ZoicCamera(**node_parameters)` plays the role of node_initialize + node_update (reference
src/zoic.cpp:1565-1720); `create_rays` is camera_create_ray (src/zoic.cpp:1752-1990) for a whole batch
of (sx, sy, lensx, lensy) samples; `stats()` returns the counters node_finish prints (:1729-1732);
`close()` is node_finish.  Parameter names, units and defaults are the reference's.

torch is used only to own device memory and streams; the data path is libzoicb's CUDA kernels.
"""
import ctypes as C
import os

import numpy as np

from . import capi

THINLENS, RAYTRACED = capi.THINLENS, capi.RAYTRACED
MODE_EXACT, MODE_GUARDED = capi.MODE_EXACT, capi.MODE_GUARDED

_BOOL_KEYS = ("useImage", "kolbSamplingLUT", "useDof")


def make_params(**kw):
    lib = capi.load()
    p = capi.Params()
    lib.zoicb_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError("unknown zoic parameter %r" % k)
        if k in ("lensDataPath", "bokehPath"):
            v = v.encode() if isinstance(v, str) else v
        elif k in _BOOL_KEYS or k == "lensModel":
            if isinstance(v, str):
                v = {"THINLENS": THINLENS, "RAYTRACED": RAYTRACED}[v]
            v = int(v)
        setattr(p, k, v)
    return p


def _constants_to_dict(c):
    n, nl = c.lensCount, c.lutSize
    out = {k: getattr(c, k) for k in ("lensCount", "apertureElement", "lutSize", "bokehWidth", "bokehHeight",
                                      "guardedSplit", "guardedInnerRetry")}
    for k in ("fov", "tan_fov", "apertureRadius", "userApertureRadius", "originShift", "apertureDistance",
              "focalLengthRatio"):
        out[k] = np.float32(getattr(c, k))
    for k in ("tracedFocalLength", "principalPlane", "focalPoint"):
        out[k] = np.array(getattr(c, k), np.float32)
    out["lenses"] = np.stack([np.array(getattr(c, k), np.float32)[:n]
                              for k in ("curvature", "thickness", "ior", "aperture", "center")], 1)
    out["lut"] = np.stack([np.array(getattr(c, k), np.float32)[:nl]
                           for k in ("lutKey", "lutMinX", "lutMinY", "lutMaxX", "lutMaxY")], 1)
    return out


def host_setup(image=None, **kw):
    """Run the host-side setup without a GPU (parity tests of the host logic).  Returns (constants, tables)."""
    lib = capi.load()
    p = make_params(**kw)
    img_ptr, w, h, nch = None, 0, 0, 0
    tabs = [None] * 4
    ptrs = [None] * 4
    if image is not None:
        img = np.ascontiguousarray(image, np.float32)
        h, w, nch = img.shape
        img_ptr = img.ctypes.data
        tabs = [np.zeros(h, np.float32), np.zeros(h, np.int32), np.zeros(h * w, np.float32), np.zeros(h * w, np.int32)]
        ptrs = [t.ctypes.data for t in tabs]
    c = capi.Constants()
    capi.check(lib.zoicb_setup_host_only(C.byref(p), img_ptr, w, h, nch, C.byref(c), *ptrs))
    return _constants_to_dict(c), (tuple(tabs) if image is not None else None)


def build_bokeh_tables(image, device=0):
    """The image-based aperture tables of `image` ([h, w, nch] floats) built on the GPU:
    returns ((cdfRow, rowIndices, cdfColumn, columnIndices), device milliseconds)."""
    lib = capi.load()
    img = np.ascontiguousarray(image, np.float32)
    h, w, nch = img.shape
    tabs = [np.zeros(h, np.float32), np.zeros(h, np.int32), np.zeros(h * w, np.float32), np.zeros(h * w, np.int32)]
    ms = C.c_float(0.0)
    capi.check(lib.zoicb_build_bokeh_tables(int(device), img.ctypes.data, w, h, nch, *[t.ctypes.data for t in tabs],
                                            C.byref(ms)))
    return tuple(tabs), ms.value


def debug_sort_orders(values):
    """(restated, library): index orders by descending value from csrc/gnu_sort.h and from the toolchain's std::sort."""
    v = np.ascontiguousarray(values, np.float32)
    a, b = np.zeros(len(v), np.int32), np.zeros(len(v), np.int32)
    capi.check(capi.load().zoicb_debug_sort_orders(v.ctypes.data, len(v), a.ctypes.data, b.ctypes.data))
    return a, b


class ZoicCamera:
    """One zoic camera node living on one CUDA device."""

    def __init__(self, image=None, device=0, mode=None, **kw):
        self.lib = capi.load()
        self.params = make_params(**kw)
        self.device = int(device)
        img_ptr, w, h, nch = None, 0, 0, 0
        if image is not None:
            self._img = np.ascontiguousarray(image, np.float32)
            h, w, nch = self._img.shape
            img_ptr = self._img.ctypes.data
        ctx = C.c_void_p()
        capi.check(self.lib.zoicb_create(C.byref(self.params), img_ptr, w, h, nch, self.device, C.byref(ctx)))
        self.ctx = ctx
        if mode is not None:
            self.set_mode(mode)

    # -- configuration -------------------------------------------------------------------------
    def set_mode(self, mode):
        capi.check(self.lib.zoicb_set_mode(self.ctx, int(mode)))

    def set_guard_scale(self, scale):
        """Validation hook: scale the decision margins of the guarded mode (1 = shipped, 0 = none)."""
        capi.check(self.lib.zoicb_set_guard_scale(self.ctx, float(scale)))

    @property
    def mode(self):
        return self.lib.zoicb_get_mode(self.ctx)

    def constants(self):
        c = capi.Constants()
        capi.check(self.lib.zoicb_get_constants(self.ctx, C.byref(c)))
        return _constants_to_dict(c)

    def bokeh_tables(self):
        h, w = self._img.shape[:2]
        tabs = [np.zeros(h, np.float32), np.zeros(h, np.int32), np.zeros(h * w, np.float32), np.zeros(h * w, np.int32)]
        capi.check(self.lib.zoicb_get_bokeh_tables(self.ctx, *[t.ctypes.data for t in tabs]))
        return tuple(tabs)

    # -- the hot path ---------------------------------------------------------------------------
    def create_rays(self, samples, seed=0, first_index=0, out=None, stream=None):
        """camera_create_ray for a CUDA tensor of samples [n, 4] -> rays [n, 8]: one 32-byte zoicb_ray per row,
        (origin.xyz, weight, dir.xyz, tries).  `split_rays` gives the two [n, 4] views."""
        import torch
        assert samples.is_cuda and samples.dtype == torch.float32 and samples.is_contiguous()
        assert samples.device.index == self.device
        n = samples.numel() // 4
        if out is None:
            out = torch.empty((n, 8), dtype=torch.float32, device=samples.device)
        assert out.is_contiguous() and out.dtype == torch.float32 and out.numel() == 8 * n
        if stream is None:
            stream = torch.cuda.current_stream(samples.device).cuda_stream
        capi.check(self.lib.zoicb_generate(self.ctx, samples.data_ptr(), n, first_index, seed, out.data_ptr(),
                                           C.c_void_p(stream)))
        return out

    def create_rays_host(self, samples, seed=0, first_index=0, out=None):
        """The same for HOST memory (numpy arrays or CPU tensors; pinned memory is copied directly): samples [n, 4]
        float32 C-contiguous, out [n, 8] float32 C-contiguous.  numpy input of another dtype / layout is converted (a copy);
        tensors and `out` must already have the right layout."""
        def is_tensor(a):
            return hasattr(a, "data_ptr")

        def check(a, what, width):
            if is_tensor(a):
                import torch
                ok = (not a.is_cuda) and a.dtype == torch.float32 and a.is_contiguous() and a.numel() % width == 0
            else:
                ok = isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags["C_CONTIGUOUS"] and a.size % width == 0
            if not ok:
                raise TypeError("%s must be a C-contiguous float32 host array with a multiple of %d elements" % (what, width))

        if not is_tensor(samples):
            samples = np.ascontiguousarray(samples, dtype=np.float32)
        check(samples, "samples", 4)

        def ptr(a):
            return a.data_ptr() if is_tensor(a) else a.ctypes.data
        n = (samples.numel() if is_tensor(samples) else samples.size) // 4
        if out is None:
            out = np.empty((n, 8), np.float32)
        check(out, "out", 8)
        if (out.numel() if is_tensor(out) else out.size) != 8 * n:
            raise ValueError("out must hold 8 floats per sample")
        capi.check(self.lib.zoicb_generate_host(self.ctx, ptr(samples), n, first_index, seed, ptr(out)))
        return out

    def create_rays_host_planar(self, samples, seed=0, first_index=0, planes=None, flags=None):
        """zoicb_generate_host_planar: host samples [n, 4] float32 in, the rays out as `planes` [6, n] float32 (origin x, y, z,
        dir x, y, z) + `flags` [n] uint8 (bits 0-6 tries, bit 7: weight == 0): 25 bytes per ray over the host link
        instead of 32.  numpy arrays or CPU tensors (pinned memory is copied directly).  Returns (planes, flags,
        live_weight); unpack_planes() rebuilds the [n, 8] records."""
        def is_tensor(a):
            return hasattr(a, "data_ptr")

        def ptr(a):
            return a.data_ptr() if is_tensor(a) else a.ctypes.data
        if not is_tensor(samples):
            samples = np.ascontiguousarray(samples, dtype=np.float32)
        elif samples.is_cuda or not samples.is_contiguous() or samples.element_size() != 4:
            raise TypeError("samples must be a C-contiguous float32 host array")
        n = (samples.numel() if is_tensor(samples) else samples.size) // 4
        if planes is None:
            planes = np.empty((6, n), np.float32)
        if flags is None:
            flags = np.empty((n,), np.uint8)
        for a, count, size, what in ((planes, 6 * n, 4, "planes"), (flags, n, 1, "flags")):
            if is_tensor(a):
                ok = (not a.is_cuda) and a.is_contiguous() and a.numel() == count and a.element_size() == size
            else:
                ok = isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"] and a.size == count and a.itemsize == size
            if not ok:
                raise TypeError("%s must be a C-contiguous host array of %d elements of %d byte(s)" % (what, count, size))
        rp = capi.RayPlanes()
        base = ptr(planes)
        for k in range(3):
            rp.origin[k] = base + 4 * n * k
            rp.dir[k] = base + 4 * n * (3 + k)
        rp.flags = ptr(flags)
        w = C.c_float(0.0)
        capi.check(self.lib.zoicb_generate_host_planar(self.ctx, ptr(samples), n, first_index, seed, C.byref(rp), C.byref(w)))
        return planes, flags, w.value

    def write_draw_file(self, path, samples, seed=0, first_index=0, indices=None):
        """draw.zoic for the reference's src/draw.py: header + the (z, y) paths of every attempt of
        the given samples ([n, 4] host array), traced on the GPU with the draw build's conventions.  `indices`
        (optional, [n] uint64) are the samples' global indices (retry streams); default first_index + i."""
        s = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1, 4)
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.uint64)
        capi.check(self.lib.zoicb_write_draw_file(self.ctx, os.fsencode(path), s.ctypes.data, s.shape[0],
                                                  None if idx is None else idx.ctypes.data, first_index, seed))

    def transform_rays(self, rays, camera_to_world, out=None, stream=None):
        """Camera -> world epilogue: rays [n, 8] on the device, camera_to_world a 3x4 row-major
        matrix (any array-like of 12 floats); returns the transformed records (in place when out is rays)."""
        import torch
        assert rays.is_cuda and rays.dtype == torch.float32 and rays.is_contiguous()
        m = np.ascontiguousarray(np.asarray(camera_to_world, dtype=np.float32).reshape(12))
        n = rays.numel() // 8
        if out is None:
            out = torch.empty_like(rays)
        if stream is None:
            stream = torch.cuda.current_stream(rays.device).cuda_stream
        capi.check(self.lib.zoicb_transform_rays(self.ctx, rays.data_ptr(), n, m.ctypes.data, out.data_ptr(),
                                                 C.c_void_p(stream)))
        return out

    def differentials(self, samples, rays, dsx, dsy, seed=0, first_index=0, out=None, stream=None):
        """Ray differentials: [n, 12] floats (dOdx, dOdy, dDdx, dDdy) for the samples [n, 4] whose
        generated records are rays [n, 8]; seed / first_index as in the create_rays call that made them."""
        import torch
        assert samples.is_cuda and rays.is_cuda and samples.is_contiguous() and rays.is_contiguous()
        assert samples.dtype == torch.float32 and rays.dtype == torch.float32
        n = samples.numel() // 4
        assert rays.numel() == 8 * n
        if out is None:
            out = torch.empty((n, 12), dtype=torch.float32, device=samples.device)
        if stream is None:
            stream = torch.cuda.current_stream(samples.device).cuda_stream
        capi.check(self.lib.zoicb_differentials(self.ctx, samples.data_ptr(), n, first_index, seed, float(dsx), float(dsy),
                                                rays.data_ptr(), out.data_ptr(), C.c_void_p(stream)))
        return out

    def transform_differentials(self, diffs, camera_to_world, out=None, stream=None):
        """Camera -> world for ray differentials [n, 12]: every vector times the 3x3 part of the 3x4 matrix."""
        import torch
        assert diffs.is_cuda and diffs.dtype == torch.float32 and diffs.is_contiguous()
        m = np.ascontiguousarray(np.asarray(camera_to_world, dtype=np.float32).reshape(12))
        n = diffs.numel() // 12
        if out is None:
            out = torch.empty_like(diffs)
        if stream is None:
            stream = torch.cuda.current_stream(diffs.device).cuda_stream
        capi.check(self.lib.zoicb_transform_differentials(self.ctx, diffs.data_ptr(), n, m.ctypes.data, out.data_ptr(),
                                                          C.c_void_p(stream)))
        return out

    def synth_samples(self, W, H, spp, seed, first_index, n, out=None, stream=None):
        """Synthetic (sx, sy, lensx, lensy) samples generated on the device."""
        import torch
        dev = torch.device("cuda", self.device)
        if out is None:
            out = torch.empty((n, 4), dtype=torch.float32, device=dev)
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        capi.check(self.lib.zoicb_synth_samples(self.ctx, W, H, spp, seed, first_index, n, out.data_ptr(),
                                                C.c_void_p(stream)))
        return out

    # -- whole-frame jobs -------------------------------------------------------------------------
    def run_job(self, W, H, spp_per_pass, sample_seed, rng_seed, first, count, tile=0, census=False, census_tol=0.0,
                windows=None, window_count=0, gather=None, gather_counts=None, serial=False):
        """zoicb_run_job: the samples [first, first + count) of a synthetic W x H frame, tile by tile through rotating
        device buffers (synthesise -> generate -> consume), optionally with the GUARDED-vs-EXACT census of every record,
        windows of records copied out for the oracle, and the NVLink gather to a consumer rank.
        Returns a dict of the result fields; with `windows` (a list of global sample indices) also "windows": a CUDA
        tensor [len(windows), window_count, 8]."""
        import torch
        job = capi.Job()
        job.W, job.H, job.spp_per_pass = int(W), int(H), int(spp_per_pass)
        job.sample_seed, job.rng_seed, job.first, job.count = int(sample_seed), int(rng_seed), int(first), int(count)
        job.tile, job.census, job.census_tol, job.serial = int(tile), int(bool(census)), float(census_tol), int(bool(serial))
        keep = []
        wout = None
        if windows:
            wf = np.ascontiguousarray(windows, dtype=np.uint64)
            wout = torch.full((len(wf), int(window_count), 8), float("nan"), dtype=torch.float32,
                              device=torch.device("cuda", self.device))
            job.n_windows, job.window_first, job.window_count, job.d_windows = len(wf), wf.ctypes.data, int(window_count), wout.data_ptr()
            keep.append(wf)
        if gather is not None:
            gc = np.ascontiguousarray(gather_counts, dtype=np.uint64)
            assert len(gc) == gather.world
            job.gather, job.gather_counts = gather.handle, gc.ctypes.data
            keep.append(gc)
        res = capi.JobResult()
        capi.check(self.lib.zoicb_run_job(self.ctx, C.byref(job), C.byref(res)))
        out = {k: getattr(res, k) for k, _ in capi.JobResult._fields_ if k not in ("stats", "census_stats")}
        out["stats"] = {k: int(getattr(res.stats, k)) for k, _ in capi.Stats._fields_}
        out["census_stats"] = {k: int(getattr(res.census_stats, k)) for k, _ in capi.Stats._fields_}
        if wout is not None:
            out["windows"] = wout
        return out

    def census(self, fast, exact, tol=1e-5, stream=None):
        """GUARDED-vs-EXACT comparison of two resident ray buffers on the device (zoicb_census)."""
        import torch
        assert fast.is_cuda and exact.is_cuda and fast.dtype == exact.dtype == torch.float32
        assert fast.is_contiguous() and exact.is_contiguous() and fast.numel() == exact.numel()
        if stream is None:
            stream = torch.cuda.current_stream(fast.device).cuda_stream
        res = capi.JobResult()
        capi.check(self.lib.zoicb_census(self.ctx, fast.data_ptr(), exact.data_ptr(), fast.numel() // 8, float(tol),
                                         C.byref(res), C.c_void_p(stream)))
        return {"rays": res.census_rays, "flips": res.census_flips, "out_of_tol": res.census_out_of_tol,
                "live": res.census_live, "max_rel_origin": res.census_max_rel_origin, "max_dir": res.census_max_dir}

    def create_times(self):
        """Milliseconds zoicb_create spent: whole call, exit-pupil LUT, image-based aperture tables."""
        t, l, b = C.c_double(0), C.c_double(0), C.c_double(0)
        capi.check(self.lib.zoicb_get_create_times(self.ctx, C.byref(t), C.byref(l), C.byref(b)))
        return {"create_ms": t.value, "lut_ms": l.value, "bokeh_ms": b.value}

    def stats(self):
        s = capi.Stats()
        capi.check(self.lib.zoicb_get_stats(self.ctx, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in capi.Stats._fields_}

    def reset_stats(self):
        capi.check(self.lib.zoicb_reset_stats(self.ctx))

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.zoicb_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Gather:
    """zoicb_gather: gather-to-consumer of ray tiles over NVLink, one process per GPU (zoic_b200/csrc/gather.cu).
    Set-up is collective: every rank creates, exports its blob, the blobs are exchanged (zoic_b200.distributed
    .connect_gather does it over torch.distributed) and every rank connects."""

    TRANSPORTS = {"fused": capi.GATHER_FUSED, "push": capi.GATHER_PUSH, "nccl": capi.GATHER_NCCL}

    def __init__(self, device, rank, world, consumer, tile_rays, slots=2, transport="fused"):
        self.lib = capi.load()
        self.rank, self.world, self.consumer, self.tile, self.slots = int(rank), int(world), int(consumer), int(tile_rays), int(slots)
        self.transport = transport
        h = C.c_void_p()
        capi.check(self.lib.zoicb_gather_create(int(device), self.rank, self.world, self.consumer, self.tile, self.slots,
                                                self.TRANSPORTS[transport], C.byref(h)))
        self.handle = h

    def export(self):
        buf = C.create_string_buffer(capi.GATHER_BLOB_BYTES)
        capi.check(self.lib.zoicb_gather_export(self.handle, buf))
        return buf.raw

    def connect(self, blobs):
        joined = b"".join(blobs)
        assert len(joined) == self.world * capi.GATHER_BLOB_BYTES
        capi.check(self.lib.zoicb_gather_connect(self.handle, joined))

    def init_nccl(self, unique_id):
        capi.check(self.lib.zoicb_gather_init_nccl(self.handle, unique_id))

    def read(self, round_, rank, offset, n):
        out = np.empty((n, 8), np.float32)
        capi.check(self.lib.zoicb_gather_read(self.handle, int(round_), int(rank), int(offset), int(n), out.ctypes.data))
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.lib.zoicb_gather_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id():
    buf = C.create_string_buffer(capi.NCCL_ID_BYTES)
    capi.check(capi.load().zoicb_nccl_unique_id(buf))
    return buf.raw


def unpack_planes(planes, flags, live_weight):
    """[n, 8] float32 records (origin, weight, dir, tries) from the planar host output of create_rays_host_planar."""
    p = np.asarray(planes).reshape(6, -1)
    f = np.asarray(flags).reshape(-1)
    out = np.empty((p.shape[1], 8), np.float32)
    out[:, 0:3] = p[0:3].T
    out[:, 3] = np.where(f & 0x80, np.float32(0.0), np.float32(live_weight))
    out[:, 4:7] = p[3:6].T
    out[:, 7] = (f & 0x7F).astype(np.float32)
    return out


def debug_lut_boxes(draws, accept, n_film, per_film, first_aperture, device=None):
    """(boxes_device or None, boxes_host): the exit-pupil LUT's bounding-box fold (reference src/zoic.cpp:1421-1440)
    on the GPU and by the host statement, for given candidate draws / accept flags."""
    d = np.ascontiguousarray(draws, np.uint32).reshape(-1)
    a = np.ascontiguousarray(accept, np.uint8).reshape(-1)
    assert len(d) == 2 * n_film * per_film and len(a) == n_film * per_film
    host = np.zeros((n_film, 4), np.float32)
    dev = np.zeros((n_film, 4), np.float32) if device is not None else None
    capi.check(capi.load().zoicb_debug_lut_boxes(int(device or 0), d.ctypes.data, a.ctypes.data, n_film, per_film,
                                                 float(first_aperture), None if dev is None else dev.ctypes.data,
                                                 host.ctypes.data))
    return dev, host


def split_rays(rays):
    """[n, 8] ray records -> (origin_w [n, 4], dir_tries [n, 4]) views."""
    return rays[:, :4], rays[:, 4:]


def kernel_launches():
    return int(capi.load().zoicb_kernel_launches())


def measure_fp32_peak(device=0):
    v = C.c_double(0.0)
    capi.check(capi.load().zoicb_measure_fp32_peak(device, C.byref(v)))
    return v.value
