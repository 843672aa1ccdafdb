"""The C-ABI shared library: it loads, it exports every symbol include/zoicb.h declares, the ctypes binding
covers all of them, and without a GPU the product path fails LOUDLY (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from zutil import ROOT

from zoic_b200 import build, capi


def _declared():
    text = open(os.path.join(ROOT, "include", "zoicb.h")).read()
    return sorted(set(re.findall(r"ZOICB_API[^;(]*?\b(zoicb_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = capi.load()
    names = _declared()
    assert len(names) >= 18
    exported = subprocess.run(["nm", "-D", "--defined-only", build.LIB], capture_output=True, text=True, check=True).stdout
    for n in names:
        assert re.search(r"\bT %s\b" % n, exported), "not exported: " + n
        assert n in capi.SYMBOLS, "no ctypes binding: " + n
        getattr(lib, n)
    assert sorted(capi.SYMBOLS) == names


def test_arnold_plugin_exports_nodeloader_only():
    out = subprocess.run(["nm", "-D", "--defined-only", build.PLUGIN], capture_output=True, text=True, check=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert syms == ["NodeLoader"]   # the reference exports exactly this symbol (src/zoic.cpp:1999)


def test_struct_layouts_match_the_header():
    """sizeof of the C structs as the C compiler sees them vs the ctypes mirrors."""
    src = r'''
    #include <stdio.h>
    #include "zoicb.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(zoicb_params), sizeof(zoicb_stats), sizeof(zoicb_constants), sizeof(zoicb_ray),
                            sizeof(zoicb_job), sizeof(zoicb_job_result), sizeof(zoicb_ray_planes), sizeof(zoicb_ray_diff)); return 0; }
    '''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        sizes = [int(x) for x in subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(capi.Params), C.sizeof(capi.Stats), C.sizeof(capi.Constants), C.sizeof(capi.Ray),
                     C.sizeof(capi.Job), C.sizeof(capi.JobResult), C.sizeof(capi.RayPlanes), 48]
    assert C.sizeof(capi.Ray) == 32


def test_defaults_are_the_reference_node_defaults():
    """reference src/zoic.cpp:1547-1562"""
    p = capi.Params()
    capi.load().zoicb_default_params(C.byref(p))
    got = {k: getattr(p, k) for k, _ in capi.Params._fields_}
    assert abs(got["sensorWidth"] - 3.6) < 1e-6 and abs(got["sensorHeight"] - 2.4) < 1e-6
    assert got["focalLength"] == 2.0 and got["fStop"] == 4.0 and got["focalDistance"] == 100.0
    assert got["useImage"] == 0 and got["lensModel"] == capi.RAYTRACED and got["kolbSamplingLUT"] == 1 and got["useDof"] == 1
    assert got["opticalVignettingDistance"] == 0.0 and got["opticalVignettingRadius"] == 1.0 and got["exposureControl"] == 0.0
    assert got["lensDataPath"] == b"" and got["bokehPath"] == b""


def test_no_gpu_fails_loudly():
    """Without a CUDA device zoicb_create must fail with ZOICB_ERR_CUDA (never fall back to the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from zoic_b200 import ZoicCamera
    with pytest.raises(capi.ZoicError) as e:
        ZoicCamera(lensModel=0, focalLength=3.5, fStop=2.8)
    assert e.value.code == capi.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_planar_host_format_unpacks_to_the_records():
    """include/zoicb.h zoicb_ray_planes: six float planes + one byte (bits 0-6 tries, bit 7: weight is 0).  A numpy
    statement of the device's packing (kernels.cu: pack_planar_kernel) against zoic_b200.unpack_planes, on records with every
    tries value, both weights, NaN and negative-zero components."""
    import numpy as np
    from zoic_b200 import unpack_planes
    rng = np.random.default_rng(3)
    n, w = 4096, np.float32(1.49)
    rec = rng.standard_normal((n, 8)).astype(np.float32)
    rec[:, 7] = rng.integers(0, 28, n).astype(np.float32)
    rec[:, 3] = np.where(rng.random(n) < 0.3, np.float32(0.0), w)
    rec[5, 0] = np.float32("nan"); rec[6, 4] = np.float32(-0.0); rec[7, 5] = np.float32("inf")
    planes = np.ascontiguousarray(rec[:, [0, 1, 2, 4, 5, 6]].T)
    flags = (rec[:, 7].astype(np.int32) & 0x7F).astype(np.uint8) | np.where(rec[:, 3] == 0, 0x80, 0).astype(np.uint8)
    back = unpack_planes(planes, flags, w)
    assert back.view(np.uint32).tolist() == rec.view(np.uint32).tolist()
    lib = capi.load()
    assert lib.zoicb_generate_host_planar(None, None, 1, 0, 0, None, None) == capi.ERR_INVALID_ARGUMENT


def test_null_arguments_are_rejected():
    lib = capi.load()
    assert lib.zoicb_create(None, None, 0, 0, 0, 0, None) == capi.ERR_INVALID_ARGUMENT
    assert lib.zoicb_generate(None, None, 1, 0, 0, None, None) == capi.ERR_INVALID_ARGUMENT
    assert lib.zoicb_get_stats(None, None) == capi.ERR_INVALID_ARGUMENT
    assert b"null" in lib.zoicb_last_error()
    assert lib.zoicb_version().startswith(b"zoicb")


def test_a_plain_c_program_links_and_calls_the_library(tmp_path):
    """include/zoicb.h + -lzoicb from C (no C++ or torch at the boundary).  Without a GPU the
    program can still read the defaults, the version, and must be refused a context with ZOICB_ERR_CUDA."""
    src = r'''
    #include <stdio.h>
    #include "zoicb.h"
    int main(void) {
        zoicb_params p; zoicb_default_params(&p);
        zoicb_ctx* cam = 0;
        p.lensModel = ZOICB_THINLENS;
        int rc = zoicb_create(&p, 0, 0, 0, 0, 0, &cam);
        printf("%s|%.1f|%d|%d|%s\n", zoicb_version(), p.focalLength, rc, cam != 0, rc ? zoicb_last_error() : "");
        if (cam) zoicb_destroy(cam);
        return 0;
    }
    '''
    c = tmp_path / "t.c"
    c.write_text(src)
    exe = tmp_path / "t"
    libdir = os.path.dirname(build.LIB)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe), "-L", libdir, "-lzoicb",
                    "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().split("|")
    assert out[0].startswith("zoicb") and out[1] == "2.0"
    import torch
    if not torch.cuda.is_available():
        assert int(out[2]) == capi.ERR_CUDA and out[3] == "0" and "no CUDA device" in out[4]
