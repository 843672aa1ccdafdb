"""Builds libzoicb.so (CUDA kernels + C ABI + Arnold-shaped adapter) in-tree for sm_100a.

nvcc cross-compiles without a GPU; the .so lands in zoic_b200/lib/ (git-ignored, but it travels with the
working tree to the GPU box).
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# ZOICB_LIBDIR: load (and build into) another directory -- A/B variants made by tools/build_variants.py
LIBDIR = os.environ.get("ZOICB_LIBDIR") or os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libzoicb.so")

SOURCES = ["capi.cu", "kernels.cu", "kolb_pool2.cu", "bokeh_build.cu", "job.cu", "gather.cu", "differentials.cu", "host_setup.cpp"]
ADAPTER = "arnold_adapter.cpp"
PLUGIN = os.path.join(LIBDIR, "libzoic_arnold.so")
HEADERS = ["camera_state.h", "lens_math.cuh", "host_setup.h", "kernels.h", "kernel_common.cuh", "gnu_sort.h",
           "capi_internal.h", "gather.h",
           os.path.join(ROOT, "include", "zoicb.h"), os.path.join(ROOT, "include", "arnold_shim", "ai.h")]

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    deps = srcs + [h for h in hdrs if os.path.exists(h)] + [os.path.abspath(__file__), os.path.join(CSRC, ADAPTER)]
    if not force and not _stale(LIB, deps) and not _stale(PLUGIN, deps):
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    # /usr/bin/g++ links libstdc++ dynamically (the /opt/gcc wrapper on PATH links it statically)
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    common = [_nvcc()] + ccbin + ARCH + [
        "-O3", "-std=c++17", "-lineinfo", "-Xptxas", "-v" if verbose else "-warn-spills",
        "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
        "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "include", "arnold_shim"),
    ] + os.environ.get("ZOICB_NVCC_FLAGS", "").split()
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = common + ["-x", "cu", "-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [_nvcc()] + ccbin + ARCH + ["-shared", "-o", LIB] + objs + ["-lpthread", "-ldl"]
    subprocess.check_call(link)
    # the Arnold-shaped plugin: NodeLoader + node callbacks on top of libzoicb.so; the Ai* host functions stay
    # undefined and are resolved by the host application (Arnold, or the test host) at load time
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-std=c++17", "-O2", "-shared", "-fPIC", "-fvisibility=hidden",
                           "-I", os.path.join(ROOT, "include", "arnold_shim"), os.path.join(CSRC, ADAPTER),
                           "-o", PLUGIN, "-L", LIBDIR, "-lzoicb", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
