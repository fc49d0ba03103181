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
    ap.add_argument("--no-configs", action="store_true", help="skip the block with every BASELINE config / shape at full size")
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


def cpu_call(mod, name, coords, w, shape=None):
    nd, shape0, _, bins, ranges, _ = WORKLOADS[name]
    shape = shape0 if shape is None else shape
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


def time_cpu(name, coords, w, shape, repeats=3):
    """The reference on one host core over the given host arrays (a slice of the timed cloud); returns the
    baseline record and the histogram of the last run (the per-bin parity check of the GPU result)."""
    mod, kind = load_cpu_impl()
    sample_n = len(coords[0])
    cpu_call(mod, name, [c[:100000] for c in coords], None if w is None else w[:100000], shape)
    best = 1e30
    h = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        h = cpu_call(mod, name, coords, w, shape)[0]
        best = min(best, time.perf_counter() - t0)
    return {"value": sample_n / best * 1e-9, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "first %.3g particles of the timed cloud, best of %d, %s, %d logical cpus on the box"
                      % (sample_n, repeats, cpu_model(), os.cpu_count()),
            "all_cores": time_cpu_all_cores(name, 4 * sample_n)}, h


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
def kernel_name(verdict, nd, shape, weighted, n, plan, fused=False):
    """The kernel of libptg_b200 that does the work for an AUTO verdict (mirrors the dispatch in
    csrc/ptg_deposit_*.cu; used to label the roofline and to look up the ncu-measured DRAM traffic)."""
    nw = 1 if shape == 0 else shape + 1
    if verdict == "tiled":
        return "k_deposit_small" if plan.cells <= 25600 else "k_deposit_tile_fx"
    if verdict == "runs":
        return "k_deposit_runs_bulk" if (nw ** nd <= 16 and not fused) else "k_deposit_runs"
    if verdict == "runs_wide":
        # odd shapes in 2-D: the widened window fed by the copy engine (ptg_deposit_runs_bulkwide.cu)
        return "k_deposit_runs_bulk_wide" if (nd == 2 and shape in (1, 3) and not fused) else "k_deposit_runs_wide"
    if shape == 0:
        return "k_deposit_global"
    return "k_deposit_quads" if (nd >= 2 and shape in (1, 2) and not fused) else "k_deposit_rows"


def launches_per_step(method, nd, shape, n, plan):
    """Kernels of libptg_b200 launched per deposit call (checked against the ncu launch lists in profiles/):
    AUTO = probe + tiled + runs (+ runs_wide where it exists) + rows (+ the merge of the staggered staging
    grid); memsets are not kernels."""
    nw = 1 if shape == 0 else shape + 1
    wide = (nw + 1) ** nd <= 27
    last = plan.dims[-1]
    staged_bytes = 2 * (plan.cells // last * ((last + 2 + 3) // 4 * 4) + 4) * 8
    stagger = shape in (1, 2) and staged_bytes <= (64 << 20) and n >= (1 << 20) and plan.cells * 8 <= (96 << 20)
    passes = 1 if plan.cells * 8 <= (96 << 20) or n < (1 << 22) else -(-plan.cells * 8 // (72 << 20))
    rows = passes * (2 if stagger else 1)
    if method in ("rows",):
        return rows
    if method in ("auto", "auto_exact"):
        if n < (1 << 21):
            return 2 if plan.cells <= 25600 else rows
        return 1 + 1 + 1 + (1 if wide else 0) + rows
    if method in ("tiled", "tiled_exact"):
        return 2
    if method == "fixed":
        return 6                      # max|w|, probe, quantum, tile kernel, direct kernel, finish
    if method == "sorted":
        passes = -(-max(plan.cells.bit_length(), 1) // 8)
        return 2 + 4 * passes + 2 * nw ** nd   # keys, offsets, 4 sort kernels per 8-bit pass, two reduce launches per colour
    return 1


def load_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (of fallback)"


def lookup_traffic(kernel, nd, shape, n):
    """ncu-measured DRAM bytes per particle of `kernel` (profiles/traffic.json, one `ncu --set full` capture per
    kernel) scaled to this launch's particle count; None when that kernel instance has not been captured."""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tr.get("%s/%dd/shape%d" % (kernel, nd, shape))
        if ent:
            return int(ent["dram_bytes_per_particle"] * n)
    except Exception:
        pass
    return None


def device_time(fn, steps, warm=3):
    """ms per call of fn(), CUDA events on the launching (current) stream after `warm` warm-up calls."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / steps


def measure_config(tag, name, n, order, method, dev, peak, shape=None, fused=False, steps=5, extra=None):
    """One line of the `configs` block: a BASELINE.json configuration (or a shape / size variant of the headline) at
    full single-GPU size, device-resident, device-timed (zero grid + deposit per step, inputs larger than L2)."""
    import torch
    from postpic_b200 import _lib, device as ptgdev
    nd, s0, _, bins, ranges, desc = WORKLOADS[name]
    shape = s0 if shape is None else shape
    coords, w = make_cloud_device(name, n, 2, order, dev)
    plan = ptgdev.Plan(list(ranges), list(bins), shape)
    grid = torch.empty(plan.dims, dtype=torch.float64, device=dev)
    bpp = BYTES_PER_PARTICLE[nd]
    fz = None
    if fused:
        me, c0 = 9.1093837015e-31, 299792458.0
        g2 = torch.Generator(device=dev).manual_seed(77)
        mom = [0.1 * me * c0 * torch.randn(n, dtype=torch.float64, device=dev, generator=g2) for _ in range(3)]
        fz = dict(kind='Ekin_MeV', px=mom[0], py=mom[1], pz=mom[2], mass=me, c=c0)
        w = torch.ones(n, dtype=torch.float64, device=dev)
        bpp += 24
    m = None if method == "auto" else method
    ms = device_time(lambda: ptgdev.deposit_into(grid, plan, coords, w, method=m, zero=True, fused=fz), steps)
    out = {"workload": name, "particles": n, "shape": shape, "order": order, "method": method, "value": n / ms * 1e-6,
           "unit": UNIT, "ms_per_step": ms, "roofline_frac": n * bpp / ms * 1e-6 / peak, "bytes_per_particle": bpp}
    if method == "auto" and n >= (1 << 21):
        try:
            v = _lib.last_probe()["mode"]
            out["kernel"] = kernel_name(v, nd, shape, True, n, plan, fused)
        except Exception:
            pass
    if fz is None:
        tot = float(w.sum())
        out["mass_rel_err"] = abs(float(grid.sum()) - tot) / tot
    if extra == "graph":
        # a launch-bound call: the same call captured in a CUDA graph (20 calls per replay) shows the device-side cost
        try:
            gph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                ptgdev.deposit_into(grid, plan, coords, w, method=m, zero=True)
            torch.cuda.current_stream().wait_stream(side)
            with torch.cuda.graph(gph):
                for _ in range(20):
                    ptgdev.deposit_into(grid, plan, coords, w, method=m, zero=True)
            ms_g = device_time(gph.replay, 20) / 20
            out["us_per_call_python_loop"] = ms * 1e3
            out["us_per_call_cuda_graph"] = ms_g * 1e3
            out["value_cuda_graph"] = n / ms_g * 1e-6
            tot = float(w.sum())
            out["mass_rel_err_graph"] = abs(float(grid.sum()) - tot) / tot
            del gph
        except Exception as exc:                                   # pragma: no cover
            out["cuda_graph_error"] = str(exc)[:160]
    if extra == "reproducible":
        # bit-reproducibility across 3 runs of the deterministic method
        ref = grid.clone()
        same = True
        for _ in range(2):
            ptgdev.deposit_into(grid, plan, coords, w, method=m, zero=True)
            same = same and bool(torch.equal(grid, ref))
        out["bit_reproducible_3_runs"] = same
    out["tag"] = tag
    del coords, w, grid, fz
    torch.cuda.empty_cache()
    return out


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
    # pinned host buffers (the e2e arm) should sit on the NUMA node of this rank's GPU: bind before allocating
    numa_cpus = parallel.bind_to_gpu_numa(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    info = _lib.device_info(local_rank)
    peak, peak_source = load_peak()

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

    probe = None
    if args.method in ("auto", "auto_exact") and n >= (1 << 21):
        try:
            probe = _lib.last_probe()
        except Exception:
            probe = None
    verdict = probe["mode"] if probe else {"global": "rows", "rows": "rows", "tiled": "tiled", "tiled_exact": "tiled",
                                           "runs": "runs", "runs_wide": "runs_wide"}.get(args.method, "rows")
    dominant = kernel_name(verdict, nd, shape, w is not None, n, plan, fused is not None)
    if args.method == "global" or shape == 0 and verdict == "rows":
        dominant = "k_deposit_global"
    if args.method == "sorted":
        dominant = "k_sorted_*"

    # the dominant kernel alone: the launch set of the method the probe chose (AUTO's gated launches of the other
    # candidates left out), K back-to-back launches between one pair of events on the launching stream
    k_method = method if method is not None else verdict
    if args.method == "auto_exact" and verdict == "tiled":
        k_method = "tiled_exact"
    durs = []
    for _ in range(3):
        grid.zero_()
        durs.append(device_time(lambda: ptgdev.deposit_into(grid, plan, coords, w, method=k_method, zero=False, fused=fused),
                                max(args.steps // 2, 3), warm=1))
    k_ms = statistics.median(durs)

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
    achieved = n * bpp / (k_ms * 1e-3) * 1e-9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": lookup_traffic(dominant, nd, shape, n), "kernel": dominant, "kernel_ms": k_ms,
                "kernel_ms_best": min(durs), "allreduce_ms": allreduce_ms,
                "allreduce_bytes": plan.cells * 8 if world > 1 else None, "bytes_per_particle": bpp,
                "peak_source": peak_source, "frac_of_nominal_8TBs": achieved / 8000.0,
                "frac_of_read_stream_7000GBs": achieved / 7000.0}

    # parity of what was just timed: mass conservation incl. ghost cells on the full run ...
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
    parity = {"mass_conservation_rel_err": abs(dep - total_w) / total_w}

    # ... and per bin against the reference on the first `cpu_sample` particles of THIS cloud (rank 0, N = 1): the
    # CPU baseline below deposits exactly those particles, so its grid is the check
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and fused is None:
        m = int(min(args.cpu_sample, n))
        hc = [c[:m].cpu().numpy() for c in coords]
        hw = w[:m].cpu().numpy() if w is not None else None
        cpu, h_ref = time_cpu(name, hc, hw, shape)
        g_s = torch.empty(plan.dims, dtype=torch.float64, device=dev)
        ptgdev.deposit_into(g_s, plan, [c[:m] for c in coords], w[:m] if w is not None else None, method=method, zero=True)
        got = g_s.cpu().numpy()
        ref = h_ref.base if h_ref.base is not None else h_ref
        wabs = float(np.abs(hw).sum()) if hw is not None else float(m)
        parity.update({"checked_particles": m, "against": cpu["kind"],
                       "per_bin_max_abs_over_max_bin": float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))),
                       "total_abs_over_sum_abs_w": float(abs(got.sum() - ref.sum()) / wabs),
                       "tolerance": 1e-12})
        parity["ok"] = bool(parity["per_bin_max_abs_over_max_bin"] <= 1e-12 and parity["total_abs_over_sum_abs_w"] <= 1e-12)
        del g_s, got, ref, h_ref, hc, hw

    # every named particle shape on this workload at N GPUs (north_star: throughput of each shape at 1/2/4/8 GPUs):
    # zero grid + deposit + all-reduce per step like `value`, max over ranks, a few steps each
    shapes_block = None
    if name == "cfg2" and args.shape is None and fused is None and not args.no_variants:
        shapes_block = {}
        for sh in (0, 1, 2, 3):
            plan_s = ptgdev.Plan(list(ranges), list(bins), sh)
            grid_s = torch.empty(plan_s.dims, dtype=torch.float64, device=dev)

            def step_s():
                ptgdev.deposit_into(grid_s, plan_s, coords, w, method=method, zero=True)
                if world > 1:
                    dist.all_reduce(grid_s, op=dist.ReduceOp.SUM)
            for _ in range(3):
                step_s()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                step_s()
            b.record()
            barrier()
            ms_s = a.elapsed_time(b) / 5
            if world > 1:
                t = torch.tensor([ms_s], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_s = float(t.item())
            shapes_block["shape%d" % sh] = {"value": n * world / ms_s * 1e-6, "unit": UNIT, "ms_per_step": ms_s,
                                            "roofline_frac_per_gpu": n * bpp / ms_s * 1e-6 / peak}
            del grid_s

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
            # N ranks: each streams ITS block through the chunked H2D pipeline (copy of chunk k+1 overlaps the
            # deposit of chunk k), one NCCL reduce to rank 0, rank 0 copies the sum back to the host
            h, _ = parallel.histogram_host_sharded(np_c, np_w, range=ranges, bins=list(bins), shape=shape, method=method,
                                                   dst=0, device_index=local_rank)
            return h
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / ksteps
        # this rank's H2D + deposit stage alone -> GB/s over its PCIe link (names the limiter when N ranks copy at once)
        g_e = torch.empty(plan.dims, dtype=torch.float64, device=dev)
        barrier()
        t0 = time.perf_counter()
        ptgdev.deposit_host_into(g_e, plan, np_c, np_w, method=method, zero=True)
        h2d_gbs = [n * bpp / (time.perf_counter() - t0) * 1e-9]
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            allv = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(allv, torch.tensor(h2d_gbs, dtype=torch.float64, device=dev))
            h2d_gbs = [float(v.item()) for v in allv]
        # where this rank's GPU hangs and where its pinned input landed (NUMA node -> pages), gathered on rank 0
        numa_diag = {"gpu_node": parallel.gpu_numa_node(local_rank), "cpus_bound": len(numa_cpus) if numa_cpus else None,
                     "pinned_pages_by_node": parallel.numa_pages_of(np_c[0].ctypes.data)}
        if world > 1:
            alld = [None] * world
            dist.all_gather_object(alld, numa_diag)
            numa_diag = alld
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
               "h2d_deposit_stage_gbs_per_rank": h2d_gbs, "numa_cpus_bound": len(numa_cpus) if numa_cpus else None, "numa": numa_diag,
               "api": "postpic_b200.histogramdd(numpy views of pinned host memory)" if world == 1 else
                      "postpic_b200.parallel.histogram_host_sharded(numpy views of pinned host memory): chunked H2D "
                      "pipeline per rank + NCCL reduce to rank 0 + grid D2H on rank 0"}
        del hc, hw, np_c, np_w, g_e

    # the same workload in the two particle orders PIC dumps come in:
    # device-timed like `value`; reported next to the headline, never instead of it
    variants = None
    configs = None
    if world == 1 and not args.no_variants and args.order == "random":
        variants = {}
        del coords, w
        torch.cuda.empty_cache()
        for order in ("sorted", "cellsorted"):
            c = measure_config(order, name, n, order, args.method, dev, peak, shape=shape, steps=args.steps)
            variants["x_sorted" if order == "sorted" else "cell_sorted"] = {
                "value": c["value"], "unit": UNIT, "ms_per_step": c["ms_per_step"], "roofline_frac": c["roofline_frac"],
                "kernel": c.get("kernel")}
    if world == 1 and not args.no_configs and args.order == "random" and name == "cfg2":
        # every BASELINE.json configuration and every particle shape at full single-GPU size (a few steps each)
        configs = []
        plan_ = [("cfg1 as named (1e6 particles: launch bound, us per call matters)", "cfg1", int(1e6), "random", "auto", None, False, 200, "graph"),
                 ("cfg1 at 1e8 particles", "cfg1", int(1e8), "random", "auto", None, False, 5, None),
                 ("cfg2 shape 0 (NGP)", "cfg2", int(1e8), "random", "auto", 0, False, 5, None),
                 ("cfg2 shape 1 (CIC)", "cfg2", int(1e8), "random", "auto", 1, False, 5, None),
                 ("cfg2 shape 3 (spline3)", "cfg2", int(1e8), "random", "auto", 3, False, 5, None),
                 ("cfg2 cell-sorted shape 0", "cfg2", int(1e8), "cellsorted", "auto", 0, False, 5, None),
                 ("cfg2 cell-sorted shape 1", "cfg2", int(1e8), "cellsorted", "auto", 1, False, 5, None),
                 ("cfg2 cell-sorted shape 3", "cfg2", int(1e8), "cellsorted", "auto", 3, False, 5, None),
                 ("cfg2 at the north-star size (1e9)", "cfg2", int(1e9), "random", "auto", None, False, 3, None),
                 ("cfg3 per-GPU share (1.25e8)", "cfg3", int(1.25e8), "random", "auto", None, False, 3, None),
                 ("cfg4 weights precomputed (5e8)", "cfg4", int(5e8), "random", "auto", None, False, 3, None),
                 ("cfg4 weight * Ekin_MeV fused (2e8)", "cfg4", int(2e8), "random", "auto", None, True, 3, None),
                 ("cfg5 AUTO (1e9)", "cfg5", int(1e9), "random", "auto", None, False, 3, None),
                 ("cfg5 deterministic SORTED (2e8)", "cfg5", int(2e8), "random", "sorted", None, False, 2, "reproducible"),
                 ("cfg5 deterministic FIXED: integer accumulation (1e9)", "cfg5", int(1e9), "random", "fixed", None, False, 3, "reproducible"),
                 ("cfg2 deterministic FIXED: integer accumulation (1e8)", "cfg2", int(1e8), "random", "fixed", None, False, 5, "reproducible"),
                 ("cfg3 deterministic FIXED: slab passes over a 275 MB accumulator (1.25e8)", "cfg3", int(1.25e8), "random", "fixed", None, False, 2, "reproducible")]
        for tag, nm, nn, order, meth, shp, fz, st, extra in plan_:
            try:
                configs.append(measure_config(tag, nm, nn, order, meth, dev, peak, shape=shp, fused=fz, steps=st, extra=extra))
            except Exception as exc:                                   # pragma: no cover
                configs.append({"tag": tag, "error": str(exc)[:200]})

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
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "variants": variants, "shapes": shapes_block,
            "configs": configs,
            "gpu_launches": args.steps * world * launches_per_step(args.method, nd, shape, n, plan),
            "clocks": clocks.summary(), "parity": parity, "auto_probe": probe,
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
