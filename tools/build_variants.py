#!/usr/bin/env python
"""Builds A/B variants of libzoicb.so that differ only in the -D flags of ONE kernel source, here in the container
(nvcc cross-compiles), so that a single GPU call can time them all:  ZOICB_LIBDIR=<dir> python bench.py ...

    python tools/build_variants.py kolb_pool2.cu name1="-DA=1 -DB=2" name2="-DA=3" ...

The other objects are taken from zoic_b200/lib/obj (run zoic_b200/build.py first).  Output: zoic_b200/lib_variants/<name>/.
"""
import os, shutil, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zoic_b200 import build as zb

def one(src, name, flags):
    out = os.path.join(ROOT, "zoic_b200", "lib_variants", name)
    os.makedirs(out, exist_ok=True)
    obj = os.path.join(out, src + ".o")
    cc = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    cmd = [zb._nvcc()] + cc + zb.ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xptxas", "-warn-spills", "-Xcompiler",
          "-fPIC,-ffp-contract=off,-fvisibility=hidden", "-I", os.path.join(ROOT, "include"), "-I",
          os.path.join(ROOT, "include", "arnold_shim")] + flags.split() + ["-x", "cu", "-c", os.path.join(zb.CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode: return name, r.stdout + r.stderr
    base = os.path.join(ROOT, "zoic_b200", "lib", "obj")
    objs = [os.path.join(base, s + ".o") for s in zb.SOURCES if s != src] + [obj]
    subprocess.check_call([zb._nvcc()] + cc + zb.ARCH + ["-shared", "-o", os.path.join(out, "libzoicb.so")] + objs + ["-lpthread", "-ldl"])
    plug = os.path.join(ROOT, "zoic_b200", "lib", "libzoic_arnold.so")
    if os.path.exists(plug): shutil.copy(plug, out)
    return name, (r.stdout + r.stderr).strip()

if __name__ == "__main__":
    src = sys.argv[1]
    jobs = [a.split("=", 1) for a in sys.argv[2:]]
    with ThreadPoolExecutor(max_workers=8) as ex:
        for name, log in ex.map(lambda j: one(src, j[0], j[1]), jobs):
            print(name, "ok" if not log else log)
