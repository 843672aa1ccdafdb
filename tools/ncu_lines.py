"""Summarise an `ncu --page source --print-source cuda,sass --csv` export: warp-level instructions executed and
stall samples per CUDA source line (top lines first), plus totals per file.  Usage: ncu_lines.py export.csv [top]"""
import csv, sys
from collections import defaultdict
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rows = list(csv.reader(open(path, errors="replace")))
cur_file = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iT = hdr.index("Thread Instructions Executed"); continue
    if hdr and r[0] != "":
        try: lines.append((cur_file, int(r[0]), r[1].strip(), int(r[iI]), int(r[iS]), int(r[iT])))
        except ValueError: pass
tot = sum(l[3] for l in lines); tots = sum(l[4] for l in lines)
print("total warp instructions %d, samples %d" % (tot, tots))
byfile = defaultdict(int)
for l in lines: byfile[l[0]] += l[3]
for f, v in sorted(byfile.items(), key=lambda x: -x[1]): print("  %-24s %5.1f%%" % (f, 100.0 * v / tot))
print("%-22s %6s %6s %5s  %s" % ("file:line", "inst%", "smpl%", "thr", "source"))
for l in sorted(lines, key=lambda x: -x[3])[:top]:
    print("%-22s %6.2f %6.2f %5.1f  %s" % ("%s:%d" % (l[0], l[1]), 100.0 * l[3] / tot, 100.0 * l[4] / max(1, tots), l[5] / max(1, l[3]), l[2][:110]))
if len(sys.argv) > 3:   # group ranges: file:lo-hi=name,...
    groups = []
    for g in sys.argv[3].split(","):
        rng, name = g.split("="); f, lh = rng.split(":"); lo, hi = lh.split("-"); groups.append((f, int(lo), int(hi), name))
    acc = defaultdict(int)
    for l in lines:
        for f, lo, hi, name in groups:
            if l[0] == f and lo <= l[1] <= hi: acc[name] += l[3]; break
        else: acc["other:" + l[0]] += l[3]
    for k, v in sorted(acc.items(), key=lambda x: -x[1]): print("  group %-28s %6.2f%%  %d" % (k, 100.0 * v / tot, v))
