#!/usr/bin/env python
"""bench.py -- particle-to-grid deposition throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload cfg2|cfg1|cfg3|cfg4|cfg5] [--order random|sorted]
                    [--method auto|global|rows|...] [--n PARTICLES]

A "step" is one complete histogram of the workload's particle cloud: zero the padded
grid, deposit every particle, and (N > 1) sum the per-rank private grids with one NCCL
all-reduce.  Default workload = BASELINE.json configs[1]: histogram2d x-px phase space,
1e8 particles, 1024x1024 bins, shape 2 (TSC), uniform synthetic cloud, float64 weights.
Weak scaling: every rank holds its own 1e8-particle cloud.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      the deposit kernel alone: algorithmic bytes (24 B/particle for 2-D weighted)
                / its CUDA-event duration, against MEASURED_PEAKS.json's hbm_gbs
  cpu_baseline  the reference's own compiled Cython module (oracle/_ref, kind "reference")
                or the C restatement (kind "port") timed on this box's host, bounded sample
  e2e           same metric through the public API with HOST (pinned) inputs: H2D + deposit
                + D2H inside the timed region
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gparticles/s deposited (2D TSC, fp64)"
UNIT = "Gparticles/s"

WORKLOADS = {
    # name: (ndim, shape, default n, bins, ranges, description)
    "cfg1": (1, 1, int(1e6), (1000,), ((0.0, 1.0),),
             "histogram1d, 1e6 particles, 1000 bins, shape=1 (CIC), float64 weights"),
    "cfg2": (2, 2, int(1e8), (1024, 1024), ((0.0, 1.0), (-1.0, 1.0)),
             "histogram2d x-px phase space, 1e8 particles, 1024x1024 bins, shape=2 (TSC), uniform synthetic cloud"),
    "cfg3": (3, 1, int(1.25e8), (256, 256, 256), ((0.0, 1.0),) * 3,
             "histogram3d density, 1.25e8 particles per GPU (1e9 over 8), 256^3 grid, shape=1"),
    "cfg4": (2, 3, int(5e8), (2048, 2048), ((0.0, 40e-6), (0.0, 40e-6)),
             "createField-like 2D energy-weighted density, 5e8 particles, 2048x2048, shape=3 (weights precomputed)"),
    "cfg5": (2, 2, int(1e9), (1024, 1024), ((0.0, 1.0), (0.0, 1.0)),
             "clustered bunch (90% of particles in 1% of cells), 2D shape=2, 1e9 particles"),
}
BYTES_PER_PARTICLE = {1: 16, 2: 24, 3: 32}     # weighted: 8 B per coordinate + 8 B weight


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--order", default="random", choices=["random", "sorted", "cellsorted"])
    ap.add_argument("--method", default="auto")
    ap.add_argument("--n", type=float, default=None, help="particles per GPU (default: the workload's)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--fused", action="store_true", help="weights = weight * Ekin_MeV(px,py,pz) evaluated inside the deposit kernel (48 B/particle in 2-D)")
    ap.add_argument("--unweighted", action="store_true", help="weights=None (the reference's separate branch)")
    ap.add_argument("--shape", type=int, default=None, help="override the workload's particle shape")
    ap.add_argument("--cpu-sample", type=float, default=2e7)
    ap.add_argument("--e2e-steps", type=int, default=None)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# synthetic clouds
# ------------------------------------------------------------------------------------------
def make_cloud_device(name, n, seed, order, device):
    """Device-resident synthetic cloud of the workload's shape."""
    import torch
    nd, shape, _, bins, ranges, _ = WORKLOADS[name]
    g = torch.Generator(device=device).manual_seed(seed)
    coords = []
    pick = None
    if name == "cfg5":
        # 90 % of the PARTICLES in the patch [0.45, 0.55)^2 = 1 % of the cells, 10 % uniform over the box
        # (SURVEY.md section 8d): one draw per particle, shared by both axes
        pick = torch.rand(n, dtype=torch.float32, device=device, generator=g) < 0.9
    for a in range(nd):
        lo, hi = ranges[a]
        u = torch.rand(n, dtype=torch.float64, device=device, generator=g)
        if pick is not None:
            u = torch.where(pick, 0.45 + 0.1 * u, u)
        coords.append(lo + (hi - lo) * u)
        del u
    del pick
    if name == "cfg4":
        # Ekin_MeV-like weights of a thermal electron species (p ~ N(0, 0.1 me c) per axis)
        p2 = sum((0.1 * torch.randn(n, dtype=torch.float64, device=device, generator=g)) ** 2 for _ in range(3))
        w = p2 / (torch.sqrt(1 + p2) + 1) * 0.51099895
        del p2
    else:
        w = torch.rand(n, dtype=torch.float64, device=device, generator=g)
    if order == "cellsorted":
        # cell-ordered, as PIC codes that sort their particles by cell write them
        key = torch.zeros(n, dtype=torch.int64, device=device)
        for a in range(nd):
            lo, hi = ranges[a]
            key = key * bins[a] + ((coords[a] - lo) / (hi - lo) * bins[a]).floor().clamp_(0, bins[a] - 1).to(torch.int64)
        idx = torch.argsort(key)
        del key
        coords = [c[idx].contiguous() for c in coords]
        w = w[idx].contiguous()
        del idx
    if order == "sorted":
        idx = torch.argsort(coords[0])
        coords = [c[idx].contiguous() for c in coords]
        w = w[idx].contiguous()
        del idx
    return coords, w


def make_cloud_host(name, n, seed):
    """The same distributions with numpy, for the CPU arms."""
    import numpy as np
    nd, shape, _, bins, ranges, _ = WORKLOADS[name]
    rng = np.random.default_rng(seed)
    coords = []
    pick = rng.random(n) < 0.9 if name == "cfg5" else None
    for a in range(nd):
        lo, hi = ranges[a]
        u = rng.random(n)
        if pick is not None:
            u = np.where(pick, 0.45 + 0.1 * u, u)
        coords.append(lo + (hi - lo) * u)
    if name == "cfg4":
        p2 = sum((0.1 * rng.standard_normal(n)) ** 2 for _ in range(3))
        w = p2 / (np.sqrt(1 + p2) + 1) * 0.51099895
    else:
        w = rng.random(n)
    return coords, w


# ------------------------------------------------------------------------------------------
# clock sampling during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler(object):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, nm in self.REASONS.items():
                    if bits & bit and nm != "gpu_idle":
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------
# CPU arms
# ------------------------------------------------------------------------------------------
def load_cpu_impl():
    """(module, kind): the reference's own compiled module if it travelled with the repo,
    else the C restatement."""
    try:
        from oracle import build_ref
        if build_ref.build(verbose=False) is not None:
            from oracle._ref import _particlestogrid as ref
            return ref, "reference"
    except Exception:
        pass
    from oracle import ptg_oracle
    ptg_oracle.lib()
    return ptg_oracle, "port"


def cpu_call(mod, name, coords, w):
    nd, shape, _, bins, ranges, _ = WORKLOADS[name]
    if nd == 1:
        return mod.histogram(coords[0], range=ranges[0], bins=bins[0], weights=w, shape=shape)
    if nd == 2:
        return mod.histogram2d(coords[0], coords[1], weights=w, range=ranges, bins=bins, shape=shape)
    return mod.histogram3d(coords[0], coords[1], coords[2], weights=w, range=ranges, bins=bins, shape=shape)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def _all_cores_worker(args):
    name, n, seed = args
    mod, _ = load_cpu_impl()
    coords, w = make_cloud_host(name, n, seed)
    t0 = time.perf_counter()
    cpu_call(mod, name, coords, w)
    return time.perf_counter() - t0


def time_cpu_all_cores(name, sample_n):
    """What the host could do with the single-threaded reference if a user ran one process per core on
    disjoint particle blocks and added the grids (the grid sum is negligible): P workers, each
    deposits sample_n / P particles; rate = sample_n / slowest worker."""
    import multiprocessing as mp
    procs = os.cpu_count() or 1
    per = max(int(sample_n // procs), 100000)
    try:
        with mp.get_context("fork").Pool(procs) as pool:
            times = pool.map(_all_cores_worker, [(name, per, 100 + k) for k in range(procs)])
        return {"value": per * procs / max(times) * 1e-9, "unit": UNIT, "cores": procs,
                "how": "%d processes x %.3g particles each, slowest worker" % (procs, per)}
    except Exception as exc:                                   # pragma: no cover
        return {"value": None, "error": str(exc)}


def time_cpu(name, sample_n, repeats=3):
    mod, kind = load_cpu_impl()
    coords, w = make_cloud_host(name, int(sample_n), 2)
    cpu_call(mod, name, [c[:100000] for c in coords], w[:100000])
    best = 1e30
    for _ in range(repeats):
        t0 = time.perf_counter()
        cpu_call(mod, name, coords, w)
        best = min(best, time.perf_counter() - t0)
    return {"value": sample_n / best * 1e-9, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "first %.3g particles of the workload, best of %d, %s, %d logical cpus on the box"
                      % (sample_n, repeats, cpu_model(), os.cpu_count()),
            "all_cores": time_cpu_all_cores(name, 4 * sample_n)}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    nd, shape, n_default, bins, ranges, desc = WORKLOADS[name]
    mod, kind = load_cpu_impl()
    sample_n = int(min(args.cpu_sample, args.n or n_default))
    coords, w = make_cloud_host(name, sample_n, 2)
    for _ in range(max(args.warmup, 1)):
        cpu_call(mod, name, coords, w)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_call(mod, name, coords, w)
    dt = (time.perf_counter() - t0) / args.steps
    value = sample_n / dt * 1e-9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "particles_per_step": sample_n, "bins": list(bins), "shape": shape,
                   "order": "random"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": "each step deposits %.3g particles of the workload (single-threaded: the "
                                   "reference holds the GIL and has no threads); %s, %d logical cpus"
                                   % (sample_n, cpu_model(), os.cpu_count()),
                         # what one process per hardware thread on disjoint particle blocks would give
                         "all_cores": time_cpu_all_cores(name, 4 * sample_n)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------
def launches_per_step(method, nd, shape, n, plan):
    """Kernels of libptg_b200 launched per deposit call (checked against the ncu launch list in
    profiles/r01/launches_cfg2_auto_final_summary.txt): AUTO = probe + tiled + runs (+ runs_wide where
    it exists) + rows (+ the merge of the staggered staging grid); memsets are not kernels."""
    nw = 1 if shape == 0 else shape + 1
    wide = (nw + 1) ** nd <= 27
    last = plan.dims[-1]
    staged_bytes = 2 * (plan.cells // last * ((last + 2 + 3) // 4 * 4) + 4) * 8
    stagger = shape in (1, 2) and staged_bytes <= (64 << 20) and n >= (1 << 20) and plan.cells * 8 <= (96 << 20)
    passes = 1 if plan.cells * 8 <= (96 << 20) or n < (1 << 22) else -(-plan.cells * 8 // (72 << 20))
    rows = passes * (2 if stagger else 1)
    if method in ("rows",):
        return rows
    if method == "auto":
        if n < (1 << 21):
            return 2 if plan.cells <= 25600 else rows
        return 1 + 1 + 1 + (1 if wide else 0) + rows
    if method == "tiled":
        return 2
    if method == "sorted":
        passes = -(-max(plan.cells.bit_length(), 1) // 8)
        return 2 + 4 * passes + 2 * nw ** nd   # keys, offsets, 4 sort kernels per 8-bit pass, two reduce launches per colour
    return 1


def run_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from postpic_b200 import _lib, device as ptgdev, parallel
    from postpic_b200.particles import _particlestogrid as ptg_mod, histogramdd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    info = _lib.device_info(local_rank)

    name = args.workload
    nd, shape, n_default, bins, ranges, desc = WORKLOADS[name]
    if args.shape is not None:
        shape = args.shape
        desc += " [shape overridden: %d]" % shape
    n = int(args.n or n_default)
    method = None if args.method == "auto" else args.method
    coords, w = make_cloud_device(name, n, 2 + rank, args.order, dev)
    if args.unweighted:
        w = None
        desc += " [unweighted]"
    plan = ptgdev.Plan(list(ranges), list(bins), shape)
    grid = torch.empty(plan.dims, dtype=torch.float64, device=dev)
    bpp = BYTES_PER_PARTICLE[nd] - (8 if args.unweighted else 0)
    fused = None
    if args.fused:
        # thermal electron momenta, p ~ N(0, (0.1 me c)^2) per axis
        me, c0 = 9.1093837015e-31, 299792458.0
        g2 = torch.Generator(device=dev).manual_seed(77 + rank)
        mom = [0.1 * me * c0 * torch.randn(n, dtype=torch.float64, device=dev, generator=g2) for _ in range(3)]
        fused = dict(kind='Ekin_MeV', px=mom[0], py=mom[1], pz=mom[2], mass=me, c=c0)
        w = torch.ones(n, dtype=torch.float64, device=dev)          # the macro-particle weight array
        bpp += 24
        desc += " [weight * Ekin_MeV fused into the deposit kernel]"

    def step():
        ptgdev.deposit_into(grid, plan, coords, w, method=method, zero=True, fused=fused)
        if world > 1:
            dist.all_reduce(grid, op=dist.ReduceOp.SUM)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = n * world / (ms_step * 1e-3) * 1e-9

    # the dominant kernel alone: events around each deposit launch on the launching stream
    durs = []
    for _ in range(args.steps):
        grid.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ptgdev.deposit_into(grid, plan, coords, w, method=method, zero=False, fused=fused)
        b.record()
        b.synchronize()
        durs.append(a.elapsed_time(b))
    k_ms = sum(durs) / len(durs)
    k_best, k_median = min(durs), statistics.median(durs)

    # the exchange step alone (N > 1): one all-reduce of the padded grid, timed like the kernel
    allreduce_ms = None
    if world > 1:
        ar = []
        for _ in range(args.steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            a.record()
            dist.all_reduce(grid, op=dist.ReduceOp.SUM)
            b.record()
            b.synchronize()
            ar.append(a.elapsed_time(b))
        t = torch.tensor([statistics.median(ar)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allreduce_ms = float(t.item())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = n * bpp / (k_ms * 1e-3) * 1e-9
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s_%s_%s" % (name, args.order, args.method)
        if key in tr and tr[key].get("n") == n:
            traffic = tr[key]["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel_ms": k_ms, "kernel_ms_best": k_best, "kernel_ms_median": k_median,
                "allreduce_ms": allreduce_ms, "allreduce_bytes": plan.cells * 8 if world > 1 else None,
                "bytes_per_particle": bpp,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "frac_of_nominal_8TBs": achieved / 8000.0}

    probe = None
    if args.method == "auto" and n >= (1 << 21):
        try:
            probe = _lib.last_probe()
        except Exception:
            probe = None

    # parity spot-check of what was just timed (cheap: mass conservation incl. ghost cells)
    if fused is not None:
        tw = ptgdev.kinetic_weights('Ekin_MeV', fused['px'], fused['py'], fused['pz'], w, mass=fused['mass'],
                                    c=fused['c']).sum()
    else:
        tw = w.sum() if w is not None else torch.tensor(float(n), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.SUM)
    total_w = float(tw)
    step()
    torch.cuda.synchronize()
    dep = float(grid.sum())
    mass_err = abs(dep - total_w) / total_w

    # end to end through the public API with host (pinned) inputs
    e2e = None
    if not args.no_e2e:
        hc = [c.cpu().pin_memory() for c in coords]
        hw = w.cpu().pin_memory() if w is not None else None
        np_c = [c.numpy() for c in hc]
        np_w = hw.numpy() if hw is not None else None
        ptg_mod.method = method
        ptg_mod.device_index = local_rank
        ksteps = args.e2e_steps or max(3, min(args.steps, 5))

        def e2e_step():
            if world == 1:
                h, edges = histogramdd(tuple(np_c), weights=np_w, range=ranges, bins=list(bins), shape=shape)
                return h
            dc = [c.to(dev, non_blocking=True) for c in hc]
            dw = hw.to(dev, non_blocking=True) if hw is not None else None
            g, _ = parallel.deposit_sharded(dc, dw, range=ranges, bins=list(bins), shape=shape, method=method)
            return g.cpu()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / ksteps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        pageable = None
        if world == 1:
            # the same call on ordinary (pageable) numpy arrays, as a postpic user would pass them
            pg_c = [np.array(c) for c in np_c]
            pg_w = np.array(np_w) if np_w is not None else None
            histogramdd(tuple(pg_c), weights=pg_w, range=ranges, bins=list(bins), shape=shape)
            t0 = time.perf_counter()
            for _ in range(3):
                histogramdd(tuple(pg_c), weights=pg_w, range=ranges, bins=list(bins), shape=shape)
            pageable = n / ((time.perf_counter() - t0) / 3) * 1e-9
            del pg_c, pg_w
        e2e = {"value": n * world / dt * 1e-9, "unit": UNIT, "h2d_bytes_per_step": n * bpp, "pageable_value": pageable,
               "d2h_bytes_per_step": plan.cells * 8, "ms_per_step": dt * 1e3, "steps": ksteps,
               "api": "postpic_b200.histogramdd(numpy views of pinned host memory)" if world == 1 else
                      "pinned host -> device copy + postpic_b200.parallel.deposit_sharded + grid D2H"}

    # the same workload in the two particle orders PIC dumps come in:
    # device-timed like `value`; reported next to the headline, never instead of it
    variants = None
    if world == 1 and not args.no_variants and args.order == "random":
        variants = {}
        del coords, w
        torch.cuda.empty_cache()
        for order in ("sorted", "cellsorted"):
            vc, vw = make_cloud_device(name, n, 2 + rank, order, dev)
            if args.unweighted:
                vw = None
            for _ in range(3):
                ptgdev.deposit_into(grid, plan, vc, vw, method=method, zero=True)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                ptgdev.deposit_into(grid, plan, vc, vw, method=method, zero=True)
            b.record()
            b.synchronize()
            ms = a.elapsed_time(b) / args.steps
            variants["x_sorted" if order == "sorted" else "cell_sorted"] = {
                "value": n / (ms * 1e-3) * 1e-9, "unit": UNIT, "ms_per_step": ms,
                "roofline_frac": n * bpp / (ms * 1e-3) * 1e-9 / peak}
            del vc, vw
            torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = time_cpu(name, int(min(args.cpu_sample, n)))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "particles_per_gpu": n, "bins": list(bins), "shape": shape,
                       "order": args.order, "method": args.method, "weights": "float64",
                       "l2": "inputs (%.1f GB per GPU) larger than L2, no flush needed" % (n * bpp * 1e-9),
                       "parallelism": "particles sharded over %d GPU(s), private grids, NCCL all-reduce" % world
                       if world > 1 else "1 GPU", "device": info["name"], "sm_count": info["sm_count"]},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "variants": variants,
            "gpu_launches": args.steps * world * launches_per_step(args.method, nd, shape, n, plan),
            "clocks": clocks.summary(),
            "parity": {"mass_conservation_rel_err": mass_err}, "auto_probe": probe,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
