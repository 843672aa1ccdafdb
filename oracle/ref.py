"""ctypes driver for the compiled UNMODIFIED reference (oracle/_ref) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The libraries are built by `make -C oracle ref` (needs /root/reference; the built .so files travel to the
GPU box, the reference sources do not).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

THINLENS, RAYTRACED = 0, 1


class RefParams(C.Structure):
    _fields_ = [
        ("sensorWidth", C.c_float), ("sensorHeight", C.c_float), ("focalLength", C.c_float),
        ("fStop", C.c_float), ("focalDistance", C.c_float),
        ("useImage", C.c_int), ("lensModel", C.c_int), ("kolbSamplingLUT", C.c_int), ("useDof", C.c_int),
        ("opticalVignettingDistance", C.c_float), ("opticalVignettingRadius", C.c_float),
        ("exposureControl", C.c_float),
        ("lensDataPath", C.c_char_p), ("bokehPath", C.c_char_p),
    ]


def available(draw=False):
    names = ["libzoic_refhost.so", "libzoic_ref_draw.so" if draw else "libzoic_ref.so"]
    return all(os.path.exists(os.path.join(REF_DIR, n)) for n in names)


_host = None


def load(draw=False):
    """Load the fake host (RTLD_GLOBAL so its xor128 interposes) and then the plugin."""
    global _host
    if _host is not None:
        return _host
    host = C.CDLL(os.path.join(REF_DIR, "libzoic_refhost.so"), mode=C.RTLD_GLOBAL)
    host.zref_open.argtypes = [C.c_char_p]
    host.zref_create.restype = C.c_void_p
    host.zref_create.argtypes = [C.POINTER(RefParams), C.c_void_p, C.c_int, C.c_int, C.c_int]
    host.zref_generate.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    host.zref_generate_one_with_state.argtypes = [C.c_void_p] * 6
    host.zref_destroy.argtypes = [C.c_void_p]
    host.zref_log.restype = C.c_char_p
    host.zref_log.argtypes = [C.c_void_p]
    host.zref_aborted.argtypes = [C.c_void_p]
    host.zref_reverse_ray.argtypes = [C.c_void_p]
    host.zref_node_name.restype = C.c_char_p
    host.zref_node_version.restype = C.c_char_p
    host.zref_sample_stream.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
    plugin = os.path.join(REF_DIR, "libzoic_ref_draw.so" if draw else "libzoic_ref.so")
    rc = host.zref_open(plugin.encode())
    if rc != 0:
        raise RuntimeError("zref_open failed: %d" % rc)
    _host = host
    return host


class RefCamera:
    """One reference camera node: NodeLoader -> Initialize -> Update done; generate() calls CreateRay per sample."""

    def __init__(self, image=None, draw=False, **kw):
        self.h = load(draw)
        p = RefParams(sensorWidth=3.6, sensorHeight=2.4, focalLength=2.0, fStop=4.0, focalDistance=100.0,
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
        img_ptr, w, hh, nch = None, 0, 0, 0
        if image is not None:
            self._img = np.ascontiguousarray(image, dtype=np.float32)
            hh, w, nch = self._img.shape
            img_ptr = self._img.ctypes.data
            if not p.bokehPath:
                p.bokehPath = b"<memory>"
        self.c = self.h.zref_create(C.byref(p), img_ptr, w, hh, nch)
        if not self.c:
            raise RuntimeError("zref_create failed")

    @property
    def log(self):
        return self.h.zref_log(self.c).decode()

    @property
    def aborted(self):
        return bool(self.h.zref_aborted(self.c))

    def generate(self, samples, seed=0, first_index=0):
        s = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1, 4)
        n = s.shape[0]
        o = np.empty((n, 4), np.float32)
        d = np.empty((n, 4), np.float32)
        st = np.zeros(3, np.uint64)
        self.h.zref_generate(self.c, s.ctypes.data, n, first_index, seed, o.ctypes.data, d.ctypes.data,
                             st.ctypes.data)
        return o, d, {"success": int(st[0]), "vignetted": int(st[1]), "attempts": int(st[2])}

    def generate_one(self, sample, state):
        s = np.asarray(sample, np.float32)
        stt = np.asarray(state, np.uint32)
        o = np.empty(4, np.float32)
        d = np.empty(4, np.float32)
        dv = np.empty(6, np.float32)
        self.h.zref_generate_one_with_state(self.c, s.ctypes.data, stt.ctypes.data, o.ctypes.data,
                                            d.ctypes.data, dv.ctypes.data)
        return o, d, dv

    def close(self):
        if self.c:
            self.h.zref_destroy(self.c)
            self.c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
