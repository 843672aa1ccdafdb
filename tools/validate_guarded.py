"""GPU validation of the GUARDED mode against the EXACT mode (which tests/test_gpu_parity.py shows to be
bit-identical to the CPU oracle): path flips, value errors against the north-star tolerance, re-run rate,
and the head-room of the decision margins (flip counts with the margins scaled down).

usage (on the GPU box): python tools/validate_guarded.py [--samples N] [--out profiles/...json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from zoic_b200 import MODE_EXACT, MODE_GUARDED, ZoicCamera, workloads  # noqa: E402


def compare(o, d, oe, de, tol=1e-5):
    flips = (o[:, 3] != oe[:, 3]) | (d[:, 3] != de[:, 3])
    live = (~flips) & (oe[:, 3] != 0)
    do = (o[:, :3] - oe[:, :3]).double().norm(dim=1)
    dd = (d[:, :3] - de[:, :3]).double().norm(dim=1)
    scale = oe[:, :3].double().norm(dim=1).clamp(min=1.0)
    rel_o = torch.where(live, do / scale, torch.zeros_like(do))
    rel_d = torch.where(live, dd, torch.zeros_like(dd))
    bad = live & ((rel_o > tol) | (rel_d > tol) | ~torch.isfinite(do) | ~torch.isfinite(dd))
    return int(flips.sum()), int(bad.sum()), float(rel_o.max()), float(rel_d.max())


def run(name, wl, n, first, scales):
    cam = ZoicCamera(image=wl.image(), **wl.params)
    s = cam.synth_samples(*wl.synth_args(), first, n)
    cam.set_mode(MODE_EXACT)
    cam.reset_stats()
    re_ = cam.create_rays(s, seed=wl.seed, first_index=first)
    torch.cuda.synchronize()
    oe, de = re_[:, :4], re_[:, 4:]
    st_e = cam.stats()
    res = {"workload": name, "samples": n, "first_index": first, "exact_stats": st_e, "scales": {}}
    cam.set_mode(MODE_GUARDED)
    for sc in scales:
        cam.set_guard_scale(sc)
        cam.reset_stats()
        r_ = cam.create_rays(s, seed=wl.seed, first_index=first)
        torch.cuda.synchronize()
        o, d = r_[:, :4], r_[:, 4:]
        st = cam.stats()
        flips, bad, eo, ed = compare(o, d, oe, de)
        same_stats = all(st[k] == st_e[k] for k in ("rays", "success", "vignetted", "attempts", "element_visits",
                                                    "total_internal_reflection"))
        res["scales"][str(sc)] = {"path_flips": flips, "out_of_tolerance": bad, "max_rel_origin_err": eo,
                                  "max_dir_err": ed, "exact_reruns": st["exact_reruns"],
                                  "rerun_fraction": st["exact_reruns"] / n, "counters_equal_exact": same_stats}
        r = res["scales"][str(sc)]
        print("%-34s first %-12d scale %-5g flips %-6d bad %-3d err_o %.2e err_d %.2e rerun %.2e stats_eq %s" % (
            name, first, sc, r["path_flips"], r["out_of_tolerance"], r["max_rel_origin_err"], r["max_dir_err"],
            r["rerun_fraction"], r["counters_equal_exact"]), flush=True)
    cam.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=1 << 26)
    ap.add_argument("--out", default="")
    ap.add_argument("--scales", default="1,0.5,0.25,0.1,0")
    a = ap.parse_args()
    scales = [float(x) for x in a.scales.split(",")]
    out = []
    cases = [("headline", workloads.headline()), ("config4", workloads.config4()), ("config3", workloads.config3())]
    cases += [("config5:" + l, workloads.config5(l)) for l in workloads.LENSES]
    for name, wl in cases:
        # three windows of the sample grid: start, middle, end (covers centre and corners of the image)
        for first in (0, (wl.n // 2 // a.samples) * a.samples, wl.n - a.samples):
            out.append(run(name, wl, a.samples, max(0, first), scales))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
