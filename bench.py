#!/usr/bin/env python
"""bench.py -- camera Mrays/s on the configurations of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload headline|config1..4] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (zoicb_generate, i.e. camera_create_ray for a whole batch) over one
batch of synthetic samples that is already resident in HBM.  The default workload is the headline metric of
BASELINE.json: the Kolb double-Gauss f/2.0 camera on a 3840x2160x256spp sample grid (2.12 G rays per GPU per
step; rank r takes pass r -- its own 256 samples of every pixel -- of a 256*N spp job: weak scaling, no data-path
collective).  Rank 0 prints ONE JSON line; see DESIGN.md section 7 for every key.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "camera Mrays/s"
UNIT = "Mrays/s"

_JSON_FD = None


def claim_stdout():
    """stdout carries the ONE JSON line and nothing else: keep the real stdout aside and point file descriptor 1 at stderr,
    so that whatever libraries print there (NCCL's version line, torchrun banners of child processes) cannot mix in."""
    global _JSON_FD
    if _JSON_FD is None:
        try:
            sys.stdout.flush()
            fd = os.dup(1)
            os.dup2(2, 1)
            _JSON_FD = fd
        except OSError:   # no usable stdout / stderr descriptors: print the ordinary way
            _JSON_FD = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
        return
    view = memoryview(data)
    while len(view):
        view = view[os.write(_JSON_FD, view):]


def flops_per_batch(model, stats):
    """Algorithmic fp32 flops (DESIGN.md section 6 / SURVEY.md 8(d)) from the kernel's exact counters."""
    if model == 1:
        return 75.0 * stats["rays"] + 56.0 * stats["attempts"] + 50.0 * stats["element_visits"]
    return 30.0 * stats["rays"] + 71.0 * stats["attempts"]


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------------
# CPU reference arm
# ------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process: build the camera with the compiled reference (or the port), time generate() on its share."""
    kind, params, use_hex, W, H, spp, seed, first, n = args
    import numpy as np
    from oracle import port
    from zoic_b200.synth import hex_bokeh_image
    image = hex_bokeh_image(255) if use_hex else None
    s = port.synth_samples(W, H, spp, seed, first, n)
    if kind == "reference":
        from oracle import ref
        cam = ref.RefCamera(image=image, **params)
    else:
        cam = port.PortCamera(image=image, **params)
    t = time.perf_counter()
    o, d, st = cam.generate(s, seed=seed, first_index=first)
    dt = time.perf_counter() - t
    cam.close()
    return n, dt, float(np.asarray(o[:, 3]).sum())


def cpu_reference_rate(wl, total_samples, cores=None):
    """Mrays/s of the reference CPU camera_create_ray on `cores` host cores, one process per core (the
    reference shares unsynchronised counters and a global RNG between threads), on a stratified sample of the
    workload: `total_samples` samples split into one contiguous slice per process, spread evenly over the
    sample grid."""
    import multiprocessing as mp
    from oracle import ref
    kind = "reference" if ref.available() else "port"
    cores = cores or os.cpu_count() or 1
    per = max(1024, total_samples // cores)
    stride = max(per, wl.n // cores)
    jobs = [(kind, wl.params, wl.image() is not None, wl.W, wl.H, wl.spp, wl.seed, min(k * stride, max(0, wl.n - per)), per)
            for k in range(cores)]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    n = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return {"value": n / slowest / 1e6, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d samples: %d contiguous slices of %d spread evenly over the %s grid; rate = samples / slowest "
                      "process (setup excluded)" % (n, cores, per, "%dx%dx%d" % (wl.W, wl.H, wl.spp)),
            "seconds": slowest, "wall_seconds": wall}


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = args.cpu_samples
    vals, last = [], None
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_rate(wl, max(per_step // 8, 8192))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = cpu_reference_rate(wl, per_step)
        vals.append(last["value"])
    wall = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": wl.describe(),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"],
                             "sample": last["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="headline")
    ap.add_argument("--mode", default="default", choices=["default", "exact", "guarded"])
    ap.add_argument("--samples", type=int, default=0, help="override samples per GPU per step (debug)")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (profiling: a small batch that still covers the whole film)")
    ap.add_argument("--e2e-samples", type=int, default=1 << 27)
    ap.add_argument("--cpu-samples", type=int, default=1 << 22, help="CPU baseline sample size (all cores)")
    ap.add_argument("--gather", action="store_true", help="N > 1: also time generation + NCCL all-gather of a 2^26-ray tile")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    claim_stdout()

    from zoic_b200 import workloads
    wl = workloads.BY_NAME[args.workload]()
    if args.spp:
        wl.spp = args.spp
        wl.name += " [spp overridden: %d]" % args.spp
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist
    from zoic_b200 import ZoicCamera, MODE_EXACT, MODE_GUARDED, camera as zcam

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    cam = ZoicCamera(image=wl.image(), device=local, **wl.params)
    if args.mode == "exact":
        cam.set_mode(MODE_EXACT)
    elif args.mode == "guarded":
        cam.set_mode(MODE_GUARDED)
    mode_name = {MODE_EXACT: "exact", MODE_GUARDED: "guarded"}[cam.mode]

    n = args.samples or wl.n
    # memory: 16 B in + 32 B out per sample resident; shrink the per-step batch if the device cannot hold it
    free, _total = torch.cuda.mem_get_info(dev)
    while n * 48 > free * 0.9:
        n //= 2
    # Weak scaling: rank r owns samples [r*n, (r+1)*n) of a W x H x spp x world job laid out PASS-major -- sample index
    # i = pass * (W*H*spp) + pixel * spp + s -- so every rank renders the whole film once (its own spp samples of every
    # pixel, its own retry streams) and all ranks carry the same mix of vignetted and clear pixels.  (Pixel-major bands
    # gave the ranks with the film's top and bottom rows 2.48 instead of 2.07 attempts per ray: 89 % efficiency at 8 GPUs
    # from load imbalance alone, profiles/r01c_bench_headline_8gpu_bands.json.)
    first = rank * n
    samples = torch.empty((n, 4), dtype=torch.float32, device=dev)
    tile = 1 << 28
    for b in range(0, n, tile):
        m = min(tile, n - b)
        cam.synth_samples(wl.W, wl.H, wl.spp, wl.seed, first + b, m, out=samples[b:b + m])   # pixel index wraps per pass
    rays = torch.empty((n, 8), dtype=torch.float32, device=dev)   # one 32-byte zoicb_ray per sample
    torch.cuda.synchronize()

    def step():
        cam.create_rays(samples, seed=wl.seed, first_index=first, out=rays)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    cam.reset_stats()
    launches0 = zcam.kernel_launches()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = zcam.kernel_launches() - launches0
    stats = cam.stats()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6

    # roofline of the (single) kernel: its launch duration is the step duration (one launch per step)
    model = wl.params["lensModel"]
    flops = flops_per_batch(model, stats) / args.steps
    kernel_s = (ms / args.steps) * 1e-3
    hbm_gbs = 48.0 * n / kernel_s / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    # DRAM traffic per launch: dram__bytes_read.sum + dram__bytes_write.sum of the same kernel from the committed
    # ncu capture (profiles/), scaled from the captured launch size to this launch (bytes per ray are size-independent)
    traffic = None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "kolb" if model == 1 else "thin"
        if key in cap:
            traffic = cap[key]["dram_bytes_per_ray"] * n
    except Exception:
        pass
    if model == 1:
        fp32_peak = zcam.measure_fp32_peak(local)
        roofline = {"bound": "fp32", "achieved": flops / kernel_s / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": flops / kernel_s / 1e12 / fp32_peak, "traffic": traffic,
                    "peak_source": "FFMA throughput measured live by zoicb_measure_fp32_peak (nominal 74.4 TFLOP/s at 1965 MHz)",
                    "flops_per_ray": flops / n, "attempts_per_ray": stats["attempts"] / max(1, stats["rays"]),
                    "element_visits_per_ray": stats["element_visits"] / max(1, stats["rays"]),
                    "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                            "peak_source": hbm_src, "bytes_per_ray": 48}}
    else:
        roofline = {"bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                    "traffic": traffic, "peak_source": hbm_src, "bytes_per_ray": 48,
                    "attempts_per_ray": stats["attempts"] / max(1, stats["rays"])}

    # end to end through the host-buffer entry point: pinned host memory in, pinned host memory out
    e2e = None
    if not args.no_e2e:
        m = min(args.e2e_samples, n)
        hs = torch.empty((m, 4), dtype=torch.float32).pin_memory()
        hr = torch.empty((m, 8), dtype=torch.float32).pin_memory()
        hs.copy_(samples[:m])
        torch.cuda.synchronize()
        cam.create_rays_host(hs, seed=wl.seed, first_index=first, out=hr)  # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            cam.create_rays_host(hs, seed=wl.seed, first_index=first, out=hr)
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * m / float(dt.item()) / 1e6, "unit": UNIT, "h2d_bytes_per_step": 16 * m,
               "d2h_bytes_per_step": 32 * m, "samples_per_step": m,
               "api": "zoicb_generate_host (pinned host buffers, 3-slot copy/compute pipeline)"}
        del hs, hr

    # optional: generation + final gather of the ray buffer over NVLink (north_star's "final NCCL gather"), on a tile
    gather = None
    if world > 1 and args.gather:
        from zoic_b200.distributed import gather_rays
        m = min(1 << 26, n)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gather_rays(rays[:m])  # warm-up
        barrier()
        g0.record()
        reps = 3
        for _ in range(reps):
            cam.create_rays(samples[:m], seed=wl.seed, first_index=first, out=rays[:m])
            gather_rays(rays[:m])
        g1.record()
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather = {"value": world * m / (float(tg.item()) * 1e-3) / 1e6, "unit": UNIT, "rays_per_rank": m,
                  "what": "generate + all-gather of the 32-byte ray records to every rank (NCCL), max over ranks"}
        # the same with double-buffered tiles: the gather of tile k travels while tile k+1 is generated
        try:
            from zoic_b200.distributed import TileGather
            m2 = min(m, n // 2)
            tiles = [rays[:m2], rays[m2:2 * m2]]
            pipe = TileGather(m2, 8, torch.float32, dev)
            reps = 6
            for k in range(2):   # warm-up
                pipe.wait(k & 1)
                cam.create_rays(samples[:m2], seed=wl.seed, first_index=first, out=tiles[k & 1])
                pipe.submit(k & 1, tiles[k & 1])
            pipe.drain()
            barrier()
            g0.record()
            for k in range(reps):
                pipe.wait(k & 1)
                cam.create_rays(samples[:m2], seed=wl.seed, first_index=first, out=tiles[k & 1])
                pipe.submit(k & 1, tiles[k & 1])
            pipe.drain()
            g1.record()
            barrier()
            tp = torch.tensor([g0.elapsed_time(g1) / reps], dtype=torch.float64, device=dev)
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
            gather["pipelined"] = {"value": world * m2 / (float(tp.item()) * 1e-3) / 1e6, "unit": UNIT, "rays_per_rank": m2,
                                   "what": "double-buffered tiles: all-gather of tile k overlaps the generation of tile k+1"}
            del pipe
        except Exception as exc:   # optional measurement: report, do not lose the bench line
            gather["pipelined"] = {"error": repr(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_reference_rate(wl, args.cpu_samples)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        one = cpu_reference_rate(wl, max(args.cpu_samples // 16, 8192), cores=1)   # the 1-thread figure too
        cpu["one_core"] = {"value": one["value"], "unit": UNIT, "sample": one["sample"]}

    if rank == 0:
        cfg = wl.describe()
        cfg.update({"samples_per_gpu_per_step": n, "arithmetic_mode": mode_name,
                    "l2": ("inputs (%.1f GB per step) are larger than L2; no flush needed" % (16.0 * n / 1e9)) if 16.0 * n > 2.6e8
                          else "batch (%.0f MB in + out) fits in L2: the number is L2-assisted" % (48.0 * n / 1e6),
                    "sharding": "rank r owns samples [r*n,(r+1)*n) of a %dx%dx%d-spp job in %d pass-major passes of %d spp: "
                                "every rank renders the whole film, no data-path collective" % (wl.W, wl.H, wl.spp * world, world, wl.spp)})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "stats": stats}
        if gather:
            line["gather"] = gather
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    cam.close()


if __name__ == "__main__":
    main()
