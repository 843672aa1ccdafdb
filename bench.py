#!/usr/bin/env python
"""bench.py -- camera Mrays/s on the configurations of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload headline|config1..4|config5:<lens>] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (camera_create_ray for every sample of the job) over one batch of synthetic
samples.  The default workload is the headline metric of BASELINE.json: the Kolb double-Gauss f/2.0 camera on a
3840x2160x256spp frame (2.12 G rays), laid out in 8 passes of 32 spp (zoic_b200/workloads.py) and STRONG-split over the
GPUs: rank r of G generates passes [r 8/G, (r+1) 8/G) = samples [r N/G, (r+1) N/G) of the one fixed job.
  * a share that fits in HBM is resident (16 B/sample in, 32 B/ray out) and a step is ONE zoicb_generate over it;
  * a share that does not (config 4: 204 GB, config 5: 1.6 TB per lens) is STREAMED by zoicb_run_job: tiles of 2^28
    samples synthesised on the device, generated, consumed by a checksum kernel, through rotating buffers.
At N > 1 the line also carries `gather`: the same job with the final gather of every ray to rank 0 over NVLink
(libzoicb's gather-to-consumer: kernels storing into rank 0's memory / copy-engine push / ncclSend-Recv), sequential and
pipelined.  Rank 0 prints ONE JSON line; see DESIGN.md section 7 for every key.
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
# CPU reference arm: the compiled unmodified reference (oracle/_ref), or the restatement where it is absent
# ------------------------------------------------------------------------------------------------------
_W = {}   # per worker process: camera + its slice of samples, built once


def _cpu_init(kind, params, use_hex, synth, seed):
    from oracle import port
    from zoic_b200.synth import hex_bokeh_image
    image = hex_bokeh_image(255) if use_hex else None
    if kind == "reference":
        from oracle import ref
        _W["cam"] = ref.RefCamera(image=image, **params)
    else:
        _W["cam"] = port.PortCamera(image=image, **params)
    _W["port"], _W["synth"], _W["seed"], _W["slices"] = port, synth, seed, {}


def _cpu_run(job):
    """One worker: generate() over its contiguous slice (synthesised once per slice and kept)."""
    first, n, threads = job
    key = (first, n)
    if key not in _W["slices"]:
        W, H, spp = _W["synth"]
        _W["slices"][key] = _W["port"].synth_samples(W, H, spp, _W["seed"], first, n)
    s = _W["slices"][key]
    t = time.perf_counter()
    if threads:
        o, d, st = _W["cam"].generate(s, seed=_W["seed"], first_index=first, nthreads=threads)
    else:
        o, d, st = _W["cam"].generate(s, seed=_W["seed"], first_index=first)
    return n, time.perf_counter() - t


class CpuReference:
    """The reference's CPU camera_create_ray on `cores` host cores: one PROCESS per core, each with its own camera node
    (the reference shares unsynchronised counters and one global RNG between the threads of a node), each on one
    contiguous slice of the workload, the slices spread evenly over the frame.  Workers and cameras persist across steps,
    so a step times generate() only."""

    def __init__(self, wl, cores=None):
        import multiprocessing as mp
        from oracle import ref
        self.wl = wl
        self.kind = "reference" if ref.available() else "port"
        self.cores = cores or os.cpu_count() or 1
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.cores, initializer=_cpu_init,
                             initargs=(self.kind, wl.params, wl.image() is not None, (wl.W, wl.H, wl.spp_per_pass), wl.seed))

    def rate(self, total_samples, threads=0):
        wl, cores = self.wl, (1 if threads else self.cores)
        per = max(1024, total_samples // cores)
        stride = max(per, wl.n // cores)
        jobs = [(min(k * stride, max(0, wl.n - per)), per, threads) for k in range(cores)]
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_run, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        n = sum(r[0] for r in res)
        slowest = max(r[1] for r in res)
        what = ("%d samples: %d contiguous slices of %d spread evenly over the %dx%dx%d frame; rate = samples / slowest "
                "process (setup and sample synthesis excluded)" % (n, cores, per, wl.W, wl.H, wl.spp))
        if threads:
            what = ("%d samples, ONE camera node shared by %d threads as shipped (unsynchronised shared counters, "
                    "src/zoic.cpp:533-534)" % (n, threads))
        return {"value": n / slowest / 1e6, "unit": UNIT, "cores": threads or cores, "kind": self.kind, "sample": what,
                "seconds": slowest, "wall_seconds": wall}

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = CpuReference(wl)
    per_step = args.cpu_samples
    for _ in range(min(args.warmup, 1)):
        cpu.rate(per_step)   # also synthesises the slices
    vals, last = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = cpu.rate(per_step)
        vals.append(last["value"])
    wall = time.perf_counter() - t0
    cpu.close()
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": wl.describe(),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"],
                             "sample": last["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def pcie_yardstick(torch, dev, barrier, reduce_max, mb=256, reps=6):
    """Raw pinned-memory copy bandwidth of this rank with ALL ranks copying at the same time: H2D alone, D2H alone and
    both directions together (GB/s; the slowest rank's figure).  The yardstick the e2e number is read against."""
    n = mb << 20
    h_up = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_up = torch.empty(n, dtype=torch.uint8, device=dev)
    d_dn = torch.empty(n, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(up, down):
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s_up):
                    d_up.copy_(h_up, non_blocking=True)
            if down:
                with torch.cuda.stream(s_dn):
                    h_dn.copy_(d_dn, non_blocking=True)
        s_up.synchronize()
        s_dn.synchronize()
        dt = reduce_max(time.perf_counter() - t0)
        return reps * n / dt / 1e9

    timed(True, True)   # warm-up
    out = {"h2d_alone": timed(True, False), "d2h_alone": timed(False, True)}
    both = timed(True, True)
    out["h2d_d2h_concurrent_each"] = both
    out["what"] = "%d MiB pinned copies, %d repetitions, all ranks at once, slowest rank; GB/s per direction" % (mb, reps)
    return out


def bind_to_gpu_numa_node(index):
    """Pins this process to the CPU cores NVML reports as local to GPU `index`, BEFORE any pinned host buffer is allocated:
    pinned pages are placed on the NUMA node of the allocating thread, so every rank's staging memory and copy-issuing
    thread end up next to its own GPU's PCIe root instead of all on node 0.  Returns the core list, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="headline")
    ap.add_argument("--mode", default="default", choices=["default", "exact", "guarded"])
    ap.add_argument("--samples", type=int, default=0, help="override samples per GPU per step (debug)")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (profiling: a small job that still covers the whole film)")
    ap.add_argument("--stream", action="store_true", help="stream the job through zoicb_run_job even if it would fit in HBM")
    ap.add_argument("--tile-log2", type=int, default=28, help="streamed jobs: samples per tile")
    ap.add_argument("--serial", action="store_true", help="streamed jobs: one stream, no overlap of synthesis / generation / consumption (A/B)")
    ap.add_argument("--e2e-samples", type=int, default=1 << 27)
    ap.add_argument("--cpu-samples", type=int, default=1 << 25, help="CPU baseline sample size (all cores)")
    ap.add_argument("--census-rays", type=int, default=2_123_366_400, help="N = 1: rays of the GUARDED-vs-EXACT census (0 = off)")
    ap.add_argument("--transports", default="fused,push,nccl", help="N > 1: gather transports to time")
    ap.add_argument("--gather-tile-log2", type=int, default=24, help="N > 1: records per rank and round of the gather")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    claim_stdout()

    from zoic_b200 import workloads
    wl = workloads.BY_NAME[args.workload]()
    if args.spp:
        wl.spp = args.spp
        wl.passes = wl.passes if args.spp % wl.passes == 0 else 1
        wl.name += " [spp overridden: %d]" % args.spp
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist
    from zoic_b200 import ZoicCamera, Gather, MODE_EXACT, MODE_GUARDED, camera as zcam
    from zoic_b200.distributed import connect_gather, job_share, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    affinity = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cam = ZoicCamera(image=wl.image(), device=local, **wl.params)
    create_times = cam.create_times()
    if args.mode == "exact":
        cam.set_mode(MODE_EXACT)
    elif args.mode == "guarded":
        cam.set_mode(MODE_GUARDED)
    mode_name = {MODE_EXACT: "exact", MODE_GUARDED: "guarded"}[cam.mode]

    # Strong split of the ONE job: whole passes per rank (every rank renders the whole film: balanced vignetting; the
    # job's rays are the same bits for every world size).  A frame whose passes the world does not divide falls back to a
    # plain contiguous split.
    try:
        first, n = job_share(wl.n, wl.passes, rank, world)
        split = "passes"
    except ValueError:
        first, n = shard_range(wl.n, rank, world)
        split = "contiguous"
    if args.samples:
        n = min(n, args.samples)
    counts = [n] * world
    W, H, spp_pp, sseed = wl.synth_args()
    free, _total = torch.cuda.mem_get_info(dev)
    resident = (not args.stream) and n * 48 <= free * 0.9
    tile = 1 << args.tile_log2
    model = wl.params["lensModel"]

    samples = rays = None
    if resident:
        samples = torch.empty((n, 4), dtype=torch.float32, device=dev)
        for b in range(0, n, 1 << 28):
            m = min(1 << 28, n - b)
            cam.synth_samples(W, H, spp_pp, sseed, first + b, m, out=samples[b:b + m])
        rays = torch.empty((n, 8), dtype=torch.float32, device=dev)   # one 32-byte zoicb_ray per sample
        torch.cuda.synchronize()

    job_results = []

    def step():
        if resident:
            cam.create_rays(samples, seed=wl.seed, first_index=first, out=rays)
        else:
            job_results.append(cam.run_job(W, H, spp_pp, sseed, wl.seed, first, n, tile=tile, serial=args.serial))

    for _ in range(args.warmup):
        step()
    barrier()
    job_results.clear()
    cam.reset_stats()
    launches0 = zcam.kernel_launches()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    clocks = sampler.stop()
    launches = zcam.kernel_launches() - launches0
    if resident:
        ms = ev0.elapsed_time(ev1)          # CUDA events on the stream the kernels were launched on
        stats = cam.stats()
        kernel_ms = ms                      # one generate call per step: its duration is the step's
        checksum = None
    else:
        ms = sum(r["device_ms"] for r in job_results)   # CUDA events inside zoicb_run_job, on its own streams
        kernel_ms = sum(r["generate_ms"] for r in job_results)
        stats = {k: sum(r["stats"][k] for r in job_results) for k in job_results[0]["stats"]}
        checksum = job_results[-1]["checksum"]
    ms_per_step = reduce_max(ms) / args.steps
    value = sum(counts) / (ms_per_step * 1e-3) / 1e6

    # roofline of the dominant kernel (the generate kernel): algorithmic flops / bytes of a launch over its own duration
    flops = flops_per_batch(model, stats) / args.steps
    kernel_s = (kernel_ms / args.steps) * 1e-3
    hbm_gbs = 48.0 * n / kernel_s / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    # DRAM traffic per launch: dram__bytes_read.sum + dram__bytes_write.sum of the same kernel from the committed ncu
    # capture (profiles/traffic.json, which names the commit it was taken at), scaled to this launch; refused (null) when
    # the kernel source is newer than the capture
    traffic, traffic_note = None, None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        # "thin" = the thin-lens retry kernel (thin_persistent_kernel: optical vignetting on); the single-attempt thin lens
        # runs another kernel (exact_kernel), for which there is no capture
        key = "kolb" if model == 1 else ("thin" if wl.params.get("opticalVignettingDistance", 0.0) > 0.0 else None)
        if key in cap:
            # the kernel's translation unit and the headers its code comes from (comma-separated in traffic.json)
            src = cap[key].get("source", "kolb_pool2.cu")
            import hashlib
            digest = hashlib.sha1(b"".join(open(os.path.join(ROOT, "zoic_b200", "csrc", f), "rb").read()
                                           for f in src.split(","))).hexdigest()[:12]
            if cap[key].get("source_sha1", digest) == digest:
                traffic = cap[key]["dram_bytes_per_ray"] * n
                traffic_note = "ncu capture %s at commit %s" % (cap[key].get("capture"), cap[key].get("commit"))
            else:
                traffic_note = "profiles/traffic.json was captured for another version of %s: not used" % src
    except Exception:
        pass
    per_launch = {"rays_per_launch": min(n, tile) if not resident else n,
                  "duration": "CUDA events around the generate launches on their stream" + ("" if resident else " (zoicb_run_job, per tile)")}
    if model == 1:
        fp32_peak = zcam.measure_fp32_peak(local)
        roofline = {"bound": "fp32", "achieved": flops / kernel_s / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": flops / kernel_s / 1e12 / fp32_peak, "traffic": traffic, "traffic_source": traffic_note,
                    "peak_source": "FFMA throughput measured live by zoicb_measure_fp32_peak (nominal 74.4 TFLOP/s at 1965 MHz); "
                                   "MEASURED_PEAKS.json carries no fp32 figure",
                    "flops_per_ray": flops / n, "attempts_per_ray": stats["attempts"] / max(1, stats["rays"]),
                    "element_visits_per_ray": stats["element_visits"] / max(1, stats["rays"]),
                    "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                            "peak_source": hbm_src, "bytes_per_ray": 48}}
    else:
        roofline = {"bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                    "traffic": traffic, "traffic_source": traffic_note, "peak_source": hbm_src, "bytes_per_ray": 48,
                    "attempts_per_ray": stats["attempts"] / max(1, stats["rays"])}
    roofline.update(per_launch)

    # end to end through the host-buffer entry point: pinned host memory in, pinned host memory out
    e2e = None
    if not args.no_e2e:
        m = min(args.e2e_samples, n)
        hs = torch.empty((m, 4), dtype=torch.float32).pin_memory()
        hout = torch.empty((32 * m,), dtype=torch.uint8).pin_memory()   # one pinned result buffer for both host formats
        hr = hout.view(torch.float32).view(m, 8)
        if resident:
            hs.copy_(samples[:m])
        else:
            hs.copy_(cam.synth_samples(W, H, spp_pp, sseed, first, m))
        torch.cuda.synchronize()
        cam.create_rays_host(hs, seed=wl.seed, first_index=first, out=hr)  # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            cam.create_rays_host(hs, seed=wl.seed, first_index=first, out=hr)
        torch.cuda.synchronize()
        dt = reduce_max((time.perf_counter() - t0) / reps)
        records = {"value": world * m / dt / 1e6, "d2h_bytes_per_step": 32 * m, "seconds_per_step": dt,
                   "api": "zoicb_generate_host (32-byte records)"}
        # the same rays as planes: 25 bytes per ray on the link instead of 32 (include/zoicb.h: zoicb_ray_planes; lossless,
        # tests/test_gpu_parity.py::test_planar_host_output_is_lossless).  The download bounds the end-to-end rate, so this
        # is the call a host that wants the rays quickly makes: the headline e2e figure.
        hp = hout[:24 * m].view(torch.float32).view(6, m)
        hf = hout[24 * m:25 * m]
        cam.create_rays_host_planar(hs, seed=wl.seed, first_index=first, planes=hp, flags=hf)   # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            cam.create_rays_host_planar(hs, seed=wl.seed, first_index=first, planes=hp, flags=hf)
        torch.cuda.synchronize()
        dtp = reduce_max((time.perf_counter() - t0) / reps)
        e2e = {"value": world * m / dtp / 1e6, "unit": UNIT, "h2d_bytes_per_step": 16 * m,
               "d2h_bytes_per_step": 25 * m, "samples_per_step": m,
               "api": "zoicb_generate_host_planar (pinned host buffers, 3-slot copy/compute pipeline; six float planes + one "
                      "byte of tries / zero-weight flag per ray)",
               "records_32B": records}
        del hs, hr, hp, hf, hout
        try:
            y = pcie_yardstick(torch, dev, barrier, reduce_max)
            per_rank_d2h = 25.0 * m / dtp / 1e9
            e2e["pcie"] = y
            e2e["d2h_gbs_per_rank"] = per_rank_d2h
            # the download (25 B/ray; 32 with records) is the binding direction; the upload (16 B/ray) shares the link
            e2e["pcie_frac"] = per_rank_d2h / y["d2h_alone"]
            records["pcie_frac"] = 32.0 * m / dt / 1e9 / y["d2h_alone"]
        except Exception as exc:   # a yardstick, not the measurement
            e2e["pcie"] = {"error": repr(exc)}

    # the whole job's GUARDED output against the EXACT mode's, every record, on the device (N = 1)
    census = None
    if world == 1 and args.census_rays and model == 1 and cam.mode == MODE_GUARDED:
        m = min(n, args.census_rays)
        t0 = time.perf_counter()
        r = cam.run_job(W, H, spp_pp, sseed, wl.seed, first, m, tile=tile, census=True)
        census = {"rays": r["census_rays"], "flips": r["census_flips"], "out_of_tol": r["census_out_of_tol"],
                  "live": r["census_live"], "max_rel_origin": r["census_max_rel_origin"], "max_dir": r["census_max_dir"],
                  "tol": 1e-5, "exact_reruns": r["stats"]["exact_reruns"], "seconds": time.perf_counter() - t0,
                  "what": "every record of the first %d samples generated in GUARDED and in EXACT mode and compared on the "
                          "device (weight and tries equal; live rays within tol)" % m}

    # the same job WITH the final gather of every ray to rank 0 over NVLink
    gather = None
    if world > 1 and not args.no_gather:
        gather = {"generate_only": value, "unit": UNIT, "consumer": 0, "records_per_rank_and_round": 1 << args.gather_tile_log2,
                  "ingest_ceiling": {"value": 900e9 / 32 * world / (world - 1) / 1e6, "unit": UNIT,
                                     "what": "NVLink 5 ingest of one GPU (900 GB/s nominal) / 32 B per ray, x G/(G-1): rank 0's own share does not travel"}}
        gtile = 1 << args.gather_tile_log2

        def gathered(transport, serial, slots, reps=2):
            g = Gather(local, rank, world, 0, gtile, slots=slots, transport=transport)
            try:
                connect_gather(g)
                best = None
                for i in range(reps + 1):
                    barrier()
                    r = cam.run_job(W, H, spp_pp, sseed, wl.seed, first, n, gather=g, gather_counts=counts, serial=serial)
                    msj = reduce_max(r["device_ms"])
                    if i and (best is None or msj < best[0]):
                        best = (msj, r)
                barrier()
                return best
            finally:
                g.close()

        # what the consumer must see: the sum of every rank's own checksum
        own = cam.run_job(W, H, spp_pp, sseed, wl.seed, first, n, tile=tile)
        parts = [None] * world
        dist.all_gather_object(parts, (own["checksum"], own["zero_weight"], own["tries_sum"]))
        expect = (sum(p[0] for p in parts) % (1 << 64), sum(p[1] for p in parts), sum(p[2] for p in parts))
        gather["expected_checksum"] = expect[0]
        runs = [("sequential", "push", True, 1)] + [(t, t, False, 3) for t in args.transports.split(",") if t]
        for name, transport, serial, slots in runs:
            try:
                msj, r = gathered(transport, serial, slots)
                entry = {"value": sum(counts) / (msj * 1e-3) / 1e6, "unit": UNIT, "ms": msj, "transport": transport,
                         "slots": slots}
                if rank == 0:
                    got = (r["checksum"], r["zero_weight"], r["tries_sum"])
                    entry["checksum_ok"] = bool(got == expect and r["consumed"] == sum(counts))
                    entry["ingest_gbs"] = 32.0 * (sum(counts) - n) / (msj * 1e-3) / 1e9
                if name == "sequential":
                    entry["what"] = "one stream, one round buffer: generate a round, push it, consume it, then the next"
                    gather["sequential"] = entry
                else:
                    gather.setdefault("pipelined", {})[name] = entry
            except Exception as exc:   # report, do not lose the bench line
                gather.setdefault("errors", {})[name] = repr(exc)
                barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        ref = CpuReference(wl)
        ref.rate(max(args.cpu_samples // 8, 8192))
        c = ref.rate(args.cpu_samples)
        cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample", "seconds")}
        ref.close()
        one = CpuReference(wl, cores=1)   # the 1-thread figure and the plugin as shipped
        o = one.rate(max(args.cpu_samples // 16, 8192))
        cpu["one_core"] = {"value": o["value"], "unit": UNIT, "sample": o["sample"]}
        if c["kind"] == "reference":
            s = one.rate(max(args.cpu_samples // 16, 8192), threads=os.cpu_count() or 1)
            cpu["as_shipped"] = {"value": s["value"], "unit": UNIT, "threads": s["cores"], "sample": s["sample"]}
        one.close()

    if rank == 0:
        run = {"samples_per_gpu_per_step": n, "first_sample_of_rank0": first, "arithmetic_mode": mode_name,
               "residency": "resident: samples and rays of the rank's share live in HBM, one zoicb_generate per step" if resident
                            else "streamed: zoicb_run_job, tiles of %d samples synthesised on the device -> generated -> consumed (checksum)" % tile,
               "l2": ("inputs (%.1f GB per step) are larger than L2; no flush needed" % (16.0 * n / 1e9)) if 16.0 * n > 2.6e8
                     else "batch (%.0f MB in + out) fits in L2: the number is L2-assisted" % (48.0 * n / 1e6),
               "sharding": "strong: rank r owns samples [r N/G, (r+1) N/G) of the one %dx%dx%d job (%s split); no data-path "
                           "collective in `value`" % (wl.W, wl.H, wl.spp, split),
               "wall_ms_per_step": reduce_max(wall_ms) / args.steps if world == 1 else None,
               "create": create_times, "cpu_affinity": ("%d cores local to the GPU: %d-%d" % (len(affinity), affinity[0], affinity[-1])) if affinity else None}
        if checksum is not None:
            run["checksum"] = checksum
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": wl.describe(), "run": run,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "stats": stats}
        if census:
            line["parity_census"] = census
        if gather:
            line["gather"] = gather
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    cam.close()


if __name__ == "__main__":
    main()
