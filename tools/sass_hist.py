#!/usr/bin/env python
"""SASS opcode histogram of one kernel of libzoicb.so (static instruction counts), with the march loop broken out.
usage: python tools/sass_hist.py [lib.so] [kernel-name-substring] > profiles/<tag>_sass_hist_<kernel>.txt"""
import re, subprocess, sys
from collections import Counter
lib = sys.argv[1] if len(sys.argv) > 1 else "zoic_b200/lib/libzoicb.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "kolb_pool2_kernelILi11ELi1ELb0ELb1ELb0ELb0"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, on, ins = None, False, []
for line in out.splitlines():
    if "Function :" in line:
        on = pat in line
        if on: name = line.split("Function :")[1].strip()
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
def op(t):
    t = t.split()[1] if t.startswith("@") else t.split()[0]
    return t
def hist(rows):
    full, base = Counter(), Counter()
    for _, t in rows:
        o = op(t); full[o] += 1; base[o.split(".")[0]] += 1
    return full, base
print("kernel:", name)
print("library:", lib)
print("static instructions:", len(ins))
full, base = hist(ins)
print("\nby opcode (static count):")
for k, v in base.most_common(): print("  %-10s %5d" % (k, v))
packed = sum(v for k, v in base.items() if k in ("FFMA2", "FMUL2", "FADD2"))
scalar = sum(v for k, v in base.items() if k in ("FFMA", "FMUL", "FADD"))
print("\npacked fp32 (FFMA2/FMUL2/FADD2): %d   scalar fp32 (FFMA/FMUL/FADD): %d   256-bit stores: %d" %
      (packed, scalar, sum(v for k, v in full.items() if k.startswith("STG") and "256" in k)))
# loops = backward branches; report the tight ones
addr2i = {a: i for i, (a, _) in enumerate(ins)}
print("\nloops (backward branches) of 20..600 instructions:")
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr2i:
        j = addr2i[int(m.group(1), 16)]
        if 20 <= i - j + 1 <= 600:
            _, b = hist(ins[j:i + 1])
            print("  %05x..%05x  %4d instructions  %s" % (ins[j][0], a, i - j + 1, dict(b.most_common(12))))
