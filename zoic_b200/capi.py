"""ctypes binding of libzoicb.so (the C ABI in include/zoicb.h).  No torch types cross this boundary."""
import ctypes as C
import os

from . import build as _build

MAX_ELEMENTS = 24
LUT_SIZE = 32

OK, ERR_INVALID_ARGUMENT, ERR_LENS_FILE, ERR_LENS_DATA, ERR_BOKEH_IMAGE, ERR_CUDA, ERR_UNSUPPORTED = range(7)
THINLENS, RAYTRACED = 0, 1
MODE_EXACT, MODE_GUARDED = 0, 1


class Params(C.Structure):
    """zoicb_params: the reference's 14 node parameters (reference src/zoic.cpp:1547-1562)."""
    _fields_ = [
        ("sensorWidth", C.c_float), ("sensorHeight", C.c_float), ("focalLength", C.c_float),
        ("fStop", C.c_float), ("focalDistance", C.c_float),
        ("useImage", C.c_int32), ("lensModel", C.c_int32), ("kolbSamplingLUT", C.c_int32), ("useDof", C.c_int32),
        ("opticalVignettingDistance", C.c_float), ("opticalVignettingRadius", C.c_float),
        ("exposureControl", C.c_float),
        ("lensDataPath", C.c_char_p), ("bokehPath", C.c_char_p),
    ]


class Ray(C.Structure):
    """zoicb_ray: one 32-byte output record."""
    _fields_ = [("origin", C.c_float * 3), ("weight", C.c_float), ("dir", C.c_float * 3), ("tries", C.c_float)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "success", "vignetted", "total_internal_reflection",
                                          "attempts", "element_visits", "exact_reruns")]


class Constants(C.Structure):
    _fields_ = [
        ("lensCount", C.c_int32), ("apertureElement", C.c_int32), ("lutSize", C.c_int32),
        ("bokehWidth", C.c_int32), ("bokehHeight", C.c_int32),
        ("fov", C.c_float), ("tan_fov", C.c_float), ("apertureRadius", C.c_float),
        ("userApertureRadius", C.c_float), ("originShift", C.c_float), ("apertureDistance", C.c_float),
        ("focalLengthRatio", C.c_float),
        ("tracedFocalLength", C.c_float * 2), ("principalPlane", C.c_float * 2), ("focalPoint", C.c_float * 2),
        ("curvature", C.c_float * MAX_ELEMENTS), ("thickness", C.c_float * MAX_ELEMENTS),
        ("ior", C.c_float * MAX_ELEMENTS), ("aperture", C.c_float * MAX_ELEMENTS),
        ("center", C.c_float * MAX_ELEMENTS),
        ("lutKey", C.c_float * LUT_SIZE), ("lutMinX", C.c_float * LUT_SIZE), ("lutMinY", C.c_float * LUT_SIZE),
        ("lutMaxX", C.c_float * LUT_SIZE), ("lutMaxY", C.c_float * LUT_SIZE),
        ("guardedSplit", C.c_int32), ("guardedInnerRetry", C.c_int32),
    ]


class Job(C.Structure):
    """zoicb_job: a range of a synthetic W x H frame, generated tile by tile (include/zoicb.h)."""
    _fields_ = [
        ("W", C.c_uint32), ("H", C.c_uint32), ("spp_per_pass", C.c_uint32), ("census", C.c_int32),
        ("sample_seed", C.c_uint64), ("rng_seed", C.c_uint64), ("first", C.c_uint64), ("count", C.c_uint64),
        ("tile", C.c_uint64), ("census_tol", C.c_float), ("n_windows", C.c_int32),
        ("window_first", C.c_void_p), ("window_count", C.c_uint64), ("d_windows", C.c_void_p),
        ("gather", C.c_void_p), ("gather_counts", C.c_void_p), ("serial", C.c_int32), ("reserved", C.c_int32),
    ]


class JobResult(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64), ("tiles", C.c_uint64), ("launches", C.c_uint64),
        ("device_ms", C.c_float), ("generate_ms", C.c_float),
        ("checksum", C.c_uint64), ("zero_weight", C.c_uint64), ("tries_sum", C.c_uint64), ("consumed", C.c_uint64),
        ("census_rays", C.c_uint64), ("census_flips", C.c_uint64), ("census_out_of_tol", C.c_uint64),
        ("census_live", C.c_uint64), ("census_max_rel_origin", C.c_float), ("census_max_dir", C.c_float),
        ("stats", Stats), ("census_stats", Stats),
    ]


GATHER_FUSED, GATHER_PUSH, GATHER_NCCL = 1, 2, 3
GATHER_BLOB_BYTES, NCCL_ID_BYTES = 192, 128

# every symbol include/zoicb.h declares: name -> (restype, argtypes)
_P = C.c_void_p


class RayPlanes(C.Structure):   # include/zoicb.h: zoicb_ray_planes
    _fields_ = [("origin", C.c_void_p * 3), ("dir", C.c_void_p * 3), ("flags", C.c_void_p)]
SYMBOLS = {
    "zoicb_default_params": (None, [C.POINTER(Params)]),
    "zoicb_create": (C.c_int, [C.POINTER(Params), _P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "zoicb_destroy": (None, [_P]),
    "zoicb_set_mode": (C.c_int, [_P, C.c_int]),
    "zoicb_get_mode": (C.c_int, [_P]),
    "zoicb_set_guard_scale": (C.c_int, [_P, C.c_float]),
    "zoicb_generate": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.c_uint64, _P, _P]),
    "zoicb_generate_host": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.c_uint64, _P]),
    "zoicb_generate_host_planar": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(RayPlanes), C.POINTER(C.c_float)]),
    "zoicb_generate_one": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P]),
    "zoicb_write_draw_file": (C.c_int, [_P, C.c_char_p, _P, C.c_uint32, _P, C.c_uint64, C.c_uint64]),
    "zoicb_transform_rays": (C.c_int, [_P, _P, C.c_uint64, _P, _P, _P]),
    "zoicb_differentials": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_float, C.c_float, _P, _P, _P]),
    "zoicb_transform_differentials": (C.c_int, [_P, _P, C.c_uint64, _P, _P, _P]),
    "zoicb_synth_samples": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, _P, _P]),
    "zoicb_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "zoicb_reset_stats": (C.c_int, [_P]),
    "zoicb_get_constants": (C.c_int, [_P, C.POINTER(Constants)]),
    "zoicb_get_bokeh_tables": (C.c_int, [_P, _P, _P, _P, _P]),
    "zoicb_setup_host_only": (C.c_int, [C.POINTER(Params), _P, C.c_int, C.c_int, C.c_int, C.POINTER(Constants), _P, _P, _P, _P]),
    "zoicb_build_bokeh_tables": (C.c_int, [C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, C.POINTER(C.c_float)]),
    "zoicb_debug_sort_orders": (C.c_int, [_P, C.c_int32, _P, _P]),
    "zoicb_measure_fp32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "zoicb_get_create_times": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "zoicb_debug_check_normalize_factor": (C.c_int, [C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "zoicb_debug_sqrt_threshold": (C.c_float, [C.c_float]),
    "zoicb_debug_lut_boxes": (C.c_int, [C.c_int, _P, _P, C.c_int32, C.c_int32, C.c_float, _P, _P]),
    "zoicb_run_job": (C.c_int, [_P, C.POINTER(Job), C.POINTER(JobResult)]),
    "zoicb_census": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_float, C.POINTER(JobResult), _P]),
    "zoicb_gather_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.POINTER(_P)]),
    "zoicb_gather_export": (C.c_int, [_P, _P]),
    "zoicb_gather_connect": (C.c_int, [_P, _P]),
    "zoicb_nccl_unique_id": (C.c_int, [_P]),
    "zoicb_gather_init_nccl": (C.c_int, [_P, _P]),
    "zoicb_gather_use_nccl_comm": (C.c_int, [_P, _P]),
    "zoicb_gather_read": (C.c_int, [_P, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, _P]),
    "zoicb_gather_destroy": (None, [_P]),
    "zoicb_kernel_launches": (C.c_uint64, []),
    "zoicb_last_error": (C.c_char_p, []),
    "zoicb_version": (C.c_char_p, []),
}

_lib = None


def library_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load libzoicb.so.  Fails loudly when the CUDA extension is missing: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        if not build_if_missing:
            raise RuntimeError("libzoicb.so is not built; run `python -m zoic_b200.build`")
        _build.build_library()
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class ZoicError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("zoicb error %d: %s" % (code, message))
        self.code = code


def check(code):
    if code != OK:
        raise ZoicError(code, load().zoicb_last_error().decode(errors="replace"))
