"""Device-timed throughput of libptg_b200 over a matrix of (workload, particle order, method) cases.

    python profiles/diag/perf_matrix.py [--n 1e8] [--steps 10] case [case ...]
    case = workload:order:method[:shape[:u]]   e.g. cfg2:random:auto  cfg2:sorted:tiled  cfg5:random:tiled_exact  cfg2:random:rows:1

Prints one line per case: Gparticles/s, fraction of MEASURED_PEAKS.json's HBM rate, the probe's verdict for
AUTO, and the worst per-bin deviation from method GLOBAL on the same input (a cheap cross-check; the parity
proper lives in tests/).  Diagnostics only -- the judged numbers come from bench.py.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload definitions and synthetic clouds)


def main():
    import torch
    from postpic_b200 import device as ptgdev, _lib
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=float, default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("cases", nargs="+")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    cache = {}
    for case in args.cases:
        parts = case.split(":")
        name, order, method = parts[0], parts[1], parts[2]
        nd, shape, n_default, bins, ranges, desc = bench.WORKLOADS[name]
        if len(parts) > 3 and parts[3] != "":
            shape = int(parts[3])
        unweighted = len(parts) > 4 and parts[4] == "u"
        n = int(args.n or n_default)
        key = (name, order, n)
        if key not in cache:
            cache.clear()
            torch.cuda.empty_cache()
            cache[key] = bench.make_cloud_device(name, n, 2, order, dev)
        coords, w = cache[key]
        if unweighted:
            w = None
        plan = ptgdev.Plan(list(ranges), list(bins), shape)
        grid = torch.empty(plan.dims, dtype=torch.float64, device=dev)
        m = None if method == "auto" else method
        for _ in range(3):
            ptgdev.deposit_into(grid, plan, coords, w, method=m, zero=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            ptgdev.deposit_into(grid, plan, coords, w, method=m, zero=True)
        b.record()
        b.synchronize()
        ms = a.elapsed_time(b) / args.steps
        bpp = 8 * nd + (0 if unweighted else 8)
        probe = ""
        if method.startswith("auto") and n >= (1 << 21):
            p = _lib.last_probe()
            probe = " probe=%s ext=%s" % (p["mode"], p["ext"][:nd])
        chk = ""
        if args.check:
            ref = torch.empty_like(grid)
            ptgdev.deposit_into(ref, plan, coords, w, method="global", zero=True)
            chk = " maxdev/max=%.2e mass=%.3e" % (float((grid - ref).abs().max()) / float(ref.abs().max()),
                                                abs(float(grid.sum()) - float(ref.sum())) / abs(float(ref.sum())))
        print("%-34s n=%.3g shape=%d %8.3f ms %8.2f Gp/s  frac=%.3f%s%s" % (
            case, n, shape, ms, n / ms * 1e-6, n * bpp / ms * 1e-6 / peak, probe, chk), flush=True)


if __name__ == "__main__":
    main()
