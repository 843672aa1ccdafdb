"""ctypes driver for the CPU restatement (oracle/zoic_port.cpp) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
`build()` compiles the restatement with g++ (seconds); it needs nothing but this repository.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libzoic_port.so")

THINLENS, RAYTRACED = 0, 1


class PortParams(C.Structure):
    _fields_ = [
        ("sensorWidth", C.c_float), ("sensorHeight", C.c_float), ("focalLength", C.c_float),
        ("fStop", C.c_float), ("focalDistance", C.c_float),
        ("useImage", C.c_int), ("lensModel", C.c_int), ("kolbSamplingLUT", C.c_int), ("useDof", C.c_int),
        ("opticalVignettingDistance", C.c_float), ("opticalVignettingRadius", C.c_float),
        ("exposureControl", C.c_float),
        ("lensDataPath", C.c_char_p), ("bokehPath", C.c_char_p),
    ]


def build(force=False):
    src = os.path.join(HERE, "zoic_port.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return LIB


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    lib = C.CDLL(build())
    lib.zport_create.restype = C.c_void_p
    lib.zport_create.argtypes = [C.POINTER(PortParams), C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.zport_destroy.argtypes = [C.c_void_p]
    lib.zport_generate.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.zport_generate_one_with_state.argtypes = [C.c_void_p] * 5
    lib.zport_synth_samples.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64,
                                        C.c_void_p]
    lib.zport_sample_stream.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
    lib.zport_transform_rays.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.zport_transform_differentials.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.zport_differentials.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_float, C.c_float,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    lib.zport_get_constants.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.zport_get_bokeh_tables.argtypes = [C.c_void_p] * 5
    _lib = lib
    return lib


def synth_samples(W, H, spp, seed, first_index, n):
    out = np.empty((n, 4), np.float32)
    load().zport_synth_samples(W, H, spp, seed, first_index, n, out.ctypes.data)
    return out


def transform_rays(rays, camera_to_world):
    """CPU statement of the camera -> world epilogue contract (include/zoicb.h: zoicb_transform_rays)."""
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    m = np.ascontiguousarray(np.asarray(camera_to_world, np.float32).reshape(12))
    out = np.empty_like(rays)
    load().zport_transform_rays(rays.ctypes.data, rays.shape[0], m.ctypes.data, out.ctypes.data)
    return out


def transform_differentials(diffs, camera_to_world):
    """CPU statement of zoicb_transform_differentials: [n, 12] differentials times the 3x3 part of the 3x4 matrix."""
    d = np.ascontiguousarray(diffs, np.float32).reshape(-1, 12)
    m = np.ascontiguousarray(np.asarray(camera_to_world, np.float32).reshape(12))
    out = np.empty_like(d)
    load().zport_transform_differentials(d.ctypes.data, d.shape[0], m.ctypes.data, out.ctypes.data)
    return out


def sample_stream(seed, index):
    out = np.empty(4, np.uint32)
    load().zport_sample_stream(seed, index, out.ctypes.data)
    return out


class PortCamera:
    """CPU restatement of one camera node (setup done in the constructor)."""

    def __init__(self, image=None, **kw):
        self.lib = load()
        p = PortParams(sensorWidth=3.6, sensorHeight=2.4, focalLength=2.0, fStop=4.0, focalDistance=100.0,
                       useImage=0, lensModel=RAYTRACED, kolbSamplingLUT=1, useDof=1,
                       opticalVignettingDistance=0.0, opticalVignettingRadius=1.0, exposureControl=0.0,
                       lensDataPath=b"", bokehPath=b"")
        for k, v in kw.items():
            if k in ("lensDataPath", "bokehPath"):
                v = v.encode() if isinstance(v, str) else v
            elif k in ("useImage", "kolbSamplingLUT", "useDof", "lensModel"):
                v = int(v)
            setattr(p, k, v)
        self.params = p
        img_ptr, w, h, nch = None, 0, 0, 0
        if image is not None:
            self._img = np.ascontiguousarray(image, dtype=np.float32)
            h, w, nch = self._img.shape
            img_ptr = self._img.ctypes.data
        err = C.c_int(0)
        self.c = self.lib.zport_create(C.byref(p), img_ptr, w, h, nch, C.byref(err))
        if not self.c:
            raise RuntimeError("zport_create failed with code %d" % err.value)

    def generate(self, samples, seed=0, first_index=0, nthreads=1):
        s = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1, 4)
        n = s.shape[0]
        o = np.empty((n, 4), np.float32)
        d = np.empty((n, 4), np.float32)
        st = np.zeros(5, np.uint64)
        self.lib.zport_generate(self.c, s.ctypes.data, n, first_index, seed, o.ctypes.data, d.ctypes.data,
                                st.ctypes.data, nthreads)
        keys = ("success", "vignetted", "attempts", "element_visits", "tir")
        return o, d, {k: int(v) for k, v in zip(keys, st)}

    def differentials(self, samples, rays, dsx, dsy, seed=0, first_index=0):
        """CPU statement of the ray-differential contract (include/zoicb.h: zoicb_differentials): [n, 12] floats
        (dOdx, dOdy, dDdx, dDdy) for samples [n, 4] whose generated records are rays [n, 8]."""
        s = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1, 4)
        r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        tries, weight = np.ascontiguousarray(r[:, 7]), np.ascontiguousarray(r[:, 3])
        out = np.empty((s.shape[0], 12), np.float32)
        self.lib.zport_differentials(self.c, s.ctypes.data, s.shape[0], first_index, seed, C.c_float(dsx), C.c_float(dsy),
                                     tries.ctypes.data, weight.ctypes.data, out.ctypes.data)
        return out

    def generate_one(self, sample, state):
        s = np.asarray(sample, np.float32)
        stt = np.asarray(state, np.uint32)
        o = np.empty(4, np.float32)
        d = np.empty(4, np.float32)
        self.lib.zport_generate_one_with_state(self.c, s.ctypes.data, stt.ctypes.data, o.ctypes.data, d.ctypes.data)
        return o, d

    def constants(self):
        sc = np.zeros(16, np.float32)
        lenses = np.zeros((16, 5), np.float32)
        lut = np.zeros((64, 6), np.float32)
        self.lib.zport_get_constants(self.c, sc.ctypes.data, lenses.ctypes.data, 16, lut.ctypes.data, 64)
        n, nl = int(sc[13]), int(sc[15])
        names = ("fov", "tan_fov", "apertureRadius", "userApertureRadius", "originShift", "apertureDistance",
                 "focalLengthRatio", "tracedFocalLength0", "tracedFocalLength1", "principalPlane0",
                 "principalPlane1", "focalPoint0", "focalPoint1")
        out = {k: np.float32(v) for k, v in zip(names, sc)}
        out.update(lensCount=n, apertureElement=int(sc[14]), lenses=lenses[:n].copy(), lut=lut[:nl, :5].copy())
        return out

    def bokeh_tables(self):
        npx = self.lib.zport_get_bokeh_tables(self.c, None, None, None, None)
        if npx == 0:
            return None
        h, w = self._img.shape[:2]
        cr = np.empty(h, np.float32); ri = np.empty(h, np.int32)
        cc = np.empty(npx, np.float32); ci = np.empty(npx, np.int32)
        self.lib.zport_get_bokeh_tables(self.c, cr.ctypes.data, ri.ctypes.data, cc.ctypes.data, ci.ctypes.data)
        return cr, ri, cc, ci

    def close(self):
        if self.c:
            self.lib.zport_destroy(self.c)
            self.c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
