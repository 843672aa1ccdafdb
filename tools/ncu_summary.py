"""Key numbers of an ncu report: python tools/ncu_summary.py report.ncu-rep  (also writes <report>_cuda_sass.csv)"""
import csv, subprocess, sys, io
from collections import defaultdict
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active"]
for w in want:
    if w in hdr: i = hdr.index(w); print("%-70s %s %s" % (w, vals[i], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
out = rep.replace(".ncu-rep", "_cuda_sass.csv"); open(out, "w").write(src)
rows = list(csv.reader(io.StringIO(src))); h = None; acc = defaultdict(int)
for r in rows:
    if not r: continue
    if r[0] == "Line No": h = r; continue
    if h and r[0].isdigit() and len(r) == len(h):
        for i, n in enumerate(h):
            if n.startswith("stall_") and "Not Issued" not in n:
                try: acc[n] += int(r[i])
                except ValueError: pass
tot = sum(acc.values())
print("stalls:", "  ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in sorted(acc.items(), key=lambda x: -x[1])[:10]))
