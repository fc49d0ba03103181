"""Every kernel family of libptg_b200 once, on small clouds, with a cross-check against method GLOBAL -- the command that
is run under `compute-sanitizer --tool memcheck` / `--tool racecheck`.  Small on purpose: the
sanitizer slows kernels down by two orders of magnitude."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from postpic_b200 import device
    rng = np.random.default_rng(0)
    ok = True
    for nd, shape, bins in ((1, 1, (200,)), (2, 2, (64, 48)), (2, 3, (700, 500)), (3, 1, (20, 18, 16))):
        n = 70000
        coords = [torch.from_numpy(rng.uniform(-0.05, 1.05, n)).cuda() for _ in range(nd)]
        w = torch.from_numpy(rng.random(n) - 0.2).cuda()
        key = torch.zeros(n, dtype=torch.int64, device='cuda')
        for a in range(nd):
            key = key * bins[a] + (coords[a] * bins[a]).floor().clamp(0, bins[a] - 1).to(torch.int64)
        order = torch.argsort(key)
        kw = dict(range=[(0, 1)] * nd if nd > 1 else (0, 1), bins=bins if nd > 1 else bins[0], shape=shape)
        for tag, cs, ws in (("random", coords, w), ("cellsorted", [c[order].contiguous() for c in coords], w[order].contiguous())):
            ref, _ = device.deposit(cs, ws, method='global', **kw)
            for m in (None, 'rows', 'tiled', 'tiled_exact', 'runs', 'runs_wide', 'sorted', 'fixed'):
                if m == 'runs_wide' and ((shape + 1) + 1) ** nd > 27:
                    continue
                got, _ = device.deposit(cs, ws, method=m, **kw)
                err = float((got - ref).abs().max()) / float(ref.abs().max())
                good = err <= 1e-12
                ok = ok and good
                print("%dd shape %d %-10s %-12s rel dev %.2e %s" % (nd, shape, tag, m, err, "ok" if good else "FAIL"), flush=True)
    torch.cuda.synchronize()
    print("SANITIZE_SMOKE", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
