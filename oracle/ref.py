"""ctypes driver for the compiled UNMODIFIED reference (oracle/_ref) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The libraries are built by `make -C oracle ref` (needs /root/reference; the built .so files travel to the
GPU box, the reference sources do not).
"""
import ctypes as C
import os

import numpy as np

import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
HOST_LIB = os.path.join(HERE, "_build", "libzoic_refhost.so")

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
    """True when the compiled reference plugin exists (built here by `make -C oracle ref`; it travels to the GPU box)."""
    return os.path.exists(os.path.join(REF_DIR, "libzoic_ref_draw.so" if draw else "libzoic_ref.so"))


def build_host():
    src = os.path.join(HERE, "ref_host.cpp")
    if not os.path.exists(HOST_LIB) or os.path.getmtime(HOST_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "host"])
    return HOST_LIB


_host = None
_plugins = {}


def load(draw=False, plugin=None):
    """Load the fake host (RTLD_GLOBAL so its xor128 interposes) and then the plugin: the compiled reference
    by default, or any other library exporting NodeLoader (zoic_b200's own Arnold plugin in its tests).
    Returns (host library, plugin index)."""
    global _host
    if _host is None:
        host = C.CDLL(build_host(), mode=C.RTLD_GLOBAL)
        host.zref_open.argtypes = [C.c_char_p]
        host.zref_create.restype = C.c_void_p
        host.zref_create.argtypes = [C.c_int, C.POINTER(RefParams), C.c_void_p, C.c_int, C.c_int, C.c_int]
        host.zref_generate.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        host.zref_generate_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        host.zref_generate_one_with_state.argtypes = [C.c_void_p] * 6
        host.zref_destroy.argtypes = [C.c_void_p]
        host.zref_log.restype = C.c_char_p
        host.zref_log.argtypes = [C.c_void_p]
        host.zref_aborted.argtypes = [C.c_void_p]
        host.zref_reverse_ray.argtypes = [C.c_void_p]
        host.zref_node_name.restype = C.c_char_p
        host.zref_node_name.argtypes = [C.c_int]
        host.zref_node_version.restype = C.c_char_p
        host.zref_node_version.argtypes = [C.c_int]
        host.zref_node_type.argtypes = [C.c_int]
        host.zref_output_type.argtypes = [C.c_int]
        host.zref_sample_stream.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
        _host = host
    if plugin is None:
        plugin = os.path.join(REF_DIR, "libzoic_ref_draw.so" if draw else "libzoic_ref.so")
    if plugin not in _plugins:
        idx = _host.zref_open(plugin.encode())
        if idx < 0:
            raise RuntimeError("zref_open(%s) failed: %d" % (plugin, idx))
        _plugins[plugin] = idx
    return _host, _plugins[plugin]


class RefCamera:
    """One reference camera node: NodeLoader -> Initialize -> Update done; generate() calls CreateRay per sample."""

    def __init__(self, image=None, draw=False, plugin=None, **kw):
        self.h, self.plugin = load(draw, plugin)
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
        self.c = self.h.zref_create(self.plugin, C.byref(p), img_ptr, w, hh, nch)
        if not self.c:
            raise RuntimeError("zref_create failed")

    @property
    def log(self):
        return self.h.zref_log(self.c).decode()

    @property
    def aborted(self):
        return bool(self.h.zref_aborted(self.c))

    def generate(self, samples, seed=0, first_index=0, nthreads=0):
        """CreateRay per sample; nthreads > 0: that many threads share this ONE node, as Arnold's render threads do
        (the reference's unsynchronised shared counters included; same rays, the retry RNG is interposed per thread)."""
        s = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1, 4)
        n = s.shape[0]
        o = np.empty((n, 4), np.float32)
        d = np.empty((n, 4), np.float32)
        st = np.zeros(3, np.uint64)
        if nthreads > 0:
            self.h.zref_generate_mt(self.c, s.ctypes.data, n, first_index, seed, o.ctypes.data, d.ctypes.data,
                                    st.ctypes.data, int(nthreads))
        else:
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
