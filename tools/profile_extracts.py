"""Turns the scratch ncu outputs of tools/collect_profiles.sh into the small, committed files under profiles/:
   python tools/profile_extracts.py <tag in gpurun_out> <name under profiles, e.g. r01_pool2>"""
import csv, io, json, os, subprocess, sys
from collections import defaultdict
tag, name = sys.argv[1], sys.argv[2]
G, P = "gpurun_out", "profiles"

# 1. launch list: one line per launch (kernel, grid, block, time, dram bytes), then shares per kernel
rows = [r for r in csv.reader(open(os.path.join(G, tag + "_launches_headline.csv"), errors="replace")) if len(r) > 14 and r[0].isdigit()]
launch = defaultdict(dict)
for r in rows:
    launch[int(r[0])]["kernel"] = r[4].split("(")[0].replace("zoicb::", "").replace("void ", "")
    launch[int(r[0])]["grid"], launch[int(r[0])]["block"] = r[8], r[7]
    launch[int(r[0])][r[12]] = float(r[14])
with open(os.path.join(P, name + "_launches_headline.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e (cold-cache, serialised: compare shares)\n")
    f.write("id,kernel,grid,block,time_ms,dram_read_MB,dram_write_MB\n")
    for i in sorted(launch):
        l = launch[i]
        f.write("%d,%s,%s,%s,%.4f,%.1f,%.1f\n" % (i, l["kernel"], l["grid"].replace(",", " "), l["block"].replace(",", " "),
                l.get("gpu__time_duration.sum", 0) / 1e6, l.get("dram__bytes_read.sum", 0) / 1e6, l.get("dram__bytes_write.sum", 0) / 1e6))
    # the bench's step launches: last two generate steps = last (main, rerun) pairs
    per = defaultdict(float)
    for i in sorted(launch):
        per[launch[i]["kernel"]] += launch[i].get("gpu__time_duration.sum", 0) / 1e6
    tot = sum(per.values())
    f.write("# share of device time per kernel over the whole command (set-up + 3 steps)\n")
    for k, v in sorted(per.items(), key=lambda x: -x[1]):
        f.write("# %-60s %10.3f ms %5.1f%%\n" % (k, v, 100 * v / tot))
main = [l for l in launch.values() if "kolb_pool" in l["kernel"]]
if main:
    m = main[-1]
    print("main kernel launch: %.3f ms, dram read %.3f GB write %.3f GB" % (m["gpu__time_duration.sum"] / 1e6, m["dram__bytes_read.sum"] / 1e9, m["dram__bytes_write.sum"] / 1e9))

# 2. raw metric extracts of the two full captures
keep = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active", "sm__warps_active", "launch__", "sm__throughput",
        "smsp__thread_inst_executed_per_inst_executed", "sm__pipe_fma", "sm__pipe_alu", "sm__pipe_fmaheavy", "sm__inst_executed_pipe_xu",
        "sm__inst_executed_pipe_lsu", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu", "dram__bytes", "dram__throughput", "gpu__dram_throughput",
        "smsp__warps_eligible", "smsp__warp_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "lts__t_sector_hit_rate",
        "smsp__sass_thread_inst_executed_op_ffma", "smsp__sass_thread_inst_executed_op_fp32", "sm__sass_thread_inst_executed_op_f", "smsp__inst_executed_op")
keep = keep + ("l1tex__t_sector", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld", "l1tex__data_pipe_lsu_wavefronts", "lts__t_sector_op_read_hit_rate")
for cap, out in ((tag + "_ncu_pool2.ncu-rep", name + "_ncu_main_raw.csv"), (tag + "_ncu_rerun.ncu-rep", name + "_ncu_rerun_raw.csv"),
                 (tag + "_ncu_thin.ncu-rep", name + "_ncu_thin_raw.csv")):
    path = os.path.join(G, cap)
    if not os.path.exists(path): continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = r[0], r[1], r[2]
    with open(os.path.join(P, out), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, python bench.py --spp 8 (config3: 32) --steps 1 --warmup 2 (a whole-film batch); kernel: %s\n" % vals[hdr.index("Kernel Name")][:120])
        f.write("metric,unit,value\n")
        for h, u, v in zip(hdr, units, vals):
            if any(h.startswith(k) for k in keep) and "realtime" not in h and ".max" not in h and ".min" not in h:
                f.write("%s,%s,%s\n" % (h, u, v))
    print("wrote", out)

# 3. traffic.json: DRAM bytes per ray of the dominant kernels, tied to the source they were captured from (bench.py refuses a
# capture whose kernel source has changed since)
import hashlib
HEADERS = "kernel_common.cuh,lens_math.cuh,camera_state.h"   # where most of both kernels' code lives
def sha(files): return hashlib.sha1(b"".join(open(os.path.join("zoic_b200/csrc", f), "rb").read() for f in files.split(","))).hexdigest()[:12]
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
traffic = {}
def per_ray(csv_path, kernel_sub, rays):
    rows = [r for r in csv.reader(open(csv_path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    by = defaultdict(dict)
    for r in rows:
        if kernel_sub in r[4]: by[int(r[0])][r[12]] = float(r[14])
    if not by: return None
    last = by[max(by)]
    return (last.get("dram__bytes_read.sum", 0) + last.get("dram__bytes_write.sum", 0)) / rays, last
hp = os.path.join(G, tag + "_launches_headline.csv")
if os.path.exists(hp):
    v = per_ray(hp, "kolb_pool2_kernel", 2123366400)
    if v:
        traffic["kolb"] = {"dram_bytes_per_ray": round(v[0], 3), "source": "kolb_pool2.cu," + HEADERS, "source_sha1": sha("kolb_pool2.cu," + HEADERS),
                           "commit": commit, "capture": "profiles/%s_launches_headline.csv" % name,
                           "what": "dram__bytes_read.sum + dram__bytes_write.sum of the last kolb_pool2_kernel launch (2,123,366,400 rays); algorithmic 48 B/ray"}
cp = os.path.join(G, tag + "_launches_config3.csv")
if os.path.exists(cp):
    v = per_ray(cp, "thin_persistent_kernel", 2123366400)   # whichever instantiation ran
    if v:
        traffic["thin"] = {"dram_bytes_per_ray": round(v[0], 3), "source": "kernels.cu," + HEADERS, "source_sha1": sha("kernels.cu," + HEADERS),
                                 "commit": commit, "capture": "profiles/%s_launches_config3.csv" % name,
                                 "what": "dram__bytes_read.sum + dram__bytes_write.sum of the last thin_persistent_kernel<1> launch (2,123,366,400 rays); algorithmic 48 B/ray"}
    rows = [r for r in csv.reader(open(cp, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    with open(os.path.join(P, name + "_launches_config3.csv"), "w") as f:
        f.write("# ncu launch list of: python bench.py --workload config3 --steps 2 --warmup 1 --no-cpu --no-e2e\nid,kernel,metric,value\n")
        for r in rows: f.write("%s,%s,%s,%s\n" % (r[0], r[4].split("(")[0].replace("zoicb::", "").replace("void ", ""), r[12], r[14]))
if traffic:
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    print("wrote traffic.json", traffic)
