"""List backward branches (loops) in one kernel's SASS with their body length and opcode histogram.
usage: sass_loops.py lib.so kernel-name-substring [min_len]"""
import re, subprocess, sys
from collections import Counter
lib, pat = sys.argv[1], sys.argv[2]; minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
on = False; ins = []
for line in out.splitlines():
    if "Function :" in line: on = pat in line; continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
print("instructions:", len(ins))
addr2i = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr2i:
        j = addr2i[int(m.group(1), 16)]
        if i - j + 1 >= minlen:
            c = Counter()
            for _, tt in ins[j:i + 1]:
                op = tt.split()[1] if tt.startswith("@") else tt.split()[0]
                c[op.split(".")[0]] += 1
            print("loop %x..%x  len %d  %s" % (ins[j][0], a, i - j + 1, dict(c.most_common())))
