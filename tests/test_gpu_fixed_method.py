"""Method FIXED (ptg_deposit_fixed.cu): deterministic and ORDER-INDEPENDENT deposition by integer accumulation
(64-bit fixed point in the shared-memory tile, 128-bit per cell in global memory).  Beyond parity with the oracle it
must be bit-identical across runs, across permutations of the particles (which also switches the kernel path the probe
picks: tile for ordered / clustered input, direct for unordered), and exact for integer weights."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

ALL = [(nd, s, w) for nd in (1, 2, 3) for s in (0, 1, 2, 3) for w in (False, True)]


def _dep(coords, w, **kw):
    from postpic_b200 import device
    g, plan = device.deposit(coords, w, method='fixed', **kw)
    return g


@pytest.mark.parametrize("nd,shape,weighted", ALL)
def test_fixed_matches_oracle(oracle, nd, shape, weighted):
    from postpic_b200.particles import _particlestogrid as ptg
    ptg.method = 'fixed'
    try:
        coords, w = cases.cloud(777 + 5 * nd + shape, 150000, nd, weighted, spill=0.15)
        hg, eg = cases.call(ptg, coords, w, shape)
        ho, eo = cases.call(oracle, coords, w, shape)
        a, b = cases.padded(hg), cases.padded(ho)
        if shape == 0 and not weighted:
            assert np.array_equal(a, b)
        else:
            assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b))
            _, asum = oracle.accepted_weight(coords, w, cases.RANGES[:nd], cases.BINS[:nd], shape)
            assert abs(a.sum() - b.sum()) <= 1e-12 * asum
    finally:
        ptg.method = None


@pytest.mark.parametrize("clustered", [False, True])
@pytest.mark.parametrize("nd,shape", [(1, 1), (2, 0), (2, 2), (2, 3), (3, 1)])
def test_fixed_is_bit_identical_across_runs_and_permutations(nd, shape, clustered):
    import torch
    n = 4_000_000
    g = torch.Generator(device='cuda').manual_seed(31 + nd + shape)
    coords = []
    pick = torch.rand(n, device='cuda', generator=g) < 0.9
    for a in range(nd):
        u = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
        coords.append(torch.where(pick, 0.45 + 0.1 * u, u * 1.1 - 0.05) if clustered else u * 1.1 - 0.05)
    w = torch.rand(n, dtype=torch.float64, device='cuda', generator=g) - 0.35        # both signs
    bins = [(4096,), (1024, 768), (96, 80, 64)][nd - 1]
    kw = dict(range=[(0, 1)] * nd if nd > 1 else (0, 1), bins=bins if nd > 1 else bins[0], shape=shape)
    g1 = _dep(coords, w, **kw)
    g2 = _dep(coords, w, **kw)
    assert torch.equal(g1, g2)
    perm = torch.randperm(n, device='cuda', generator=g)
    g3 = _dep([c[perm] for c in coords], w[perm], **kw)
    assert torch.equal(g1, g3)                       # any order of the particles: the same bits
    order = torch.argsort(coords[0])                 # x-ordered: the probe sends this to the shared-memory tile
    g4 = _dep([c[order].contiguous() for c in coords], w[order].contiguous(), **kw)
    assert torch.equal(g1, g4)
    from postpic_b200 import device
    ref, _ = device.deposit(coords, w, method='global', **kw)
    assert float((g1 - ref).abs().max()) <= 1e-12 * float(ref.abs().max())


def test_fixed_integer_weights_exact_and_heavy_cell():
    import torch
    from postpic_b200 import device
    n = 30_000_000
    g = torch.Generator(device='cuda').manual_seed(5)
    x = 0.5 + 1e-4 * torch.rand(n, dtype=torch.float64, device='cuda', generator=g)          # all in one cell
    y = 0.5 + 1e-4 * torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    w = torch.randint(-3, 9, (n,), device='cuda', generator=g).to(torch.float64)
    kw = dict(range=((0, 1), (0, 1)), bins=(1024, 1024))
    got = _dep([x, y], w, shape=0, **kw)
    assert float(got.sum()) == float(w.sum()) and float(got.abs().max()) == abs(float(w.sum()))
    got = _dep([x, y], w, shape=2, **kw)                                               # 3e7 particles on 9 cells
    ref, _ = device.deposit([x, y], w, method='global', shape=2, **kw)
    assert float((got - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    # unordered over the whole grid (direct 128-bit path), integer weights, shape 0: exact
    x2 = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    y2 = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    got = _dep([x2, y2], w, shape=0, **kw)
    ref, _ = device.deposit([x2, y2], w, method='global', shape=0, **kw)
    assert torch.equal(got, ref)


def test_fixed_nonfinite_weights_and_dynamic_range():
    import torch
    from postpic_b200 import device
    rng = np.random.default_rng(8)
    n = 2_000_000
    x, y = rng.random(n), rng.random(n)
    w = 10.0 ** rng.uniform(-9, 4, n) * rng.choice([-1.0, 1.0], n)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    kw = dict(range=((0, 1), (0, 1)), bins=(96, 64), shape=2)
    ref, _ = device.deposit([dx, dy], torch.from_numpy(w).cuda(), method='global', **kw)
    got = _dep([dx, dy], torch.from_numpy(w).cuda(), **kw)
    assert float((got - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    w[1234] = np.inf
    w[777777] = np.nan
    dw = torch.from_numpy(w).cuda()
    ref, _ = device.deposit([dx, dy], dw, method='global', **kw)
    got = _dep([dx, dy], dw, **kw)
    assert torch.equal(torch.isfinite(got), torch.isfinite(ref))
    fin = torch.isfinite(ref)
    assert float((got[fin] - ref[fin]).abs().max()) <= 1e-12 * float(ref[fin].abs().max())
    assert torch.equal(got, _dep([dx, dy], dw, **kw)) or bool(torch.isnan(got).any())


def test_fixed_through_the_numpy_drop_in_is_reproducible():
    """Several pipeline chunks (host arrays): the chunks are deposited in order on one stream, so the result is
    reproducible there too."""
    from postpic_b200.particles import _particlestogrid as ptg
    rng = np.random.default_rng(3)
    n = 9_000_000
    x, y, w = rng.random(n), rng.random(n), rng.random(n)
    ptg.method = 'fixed'
    try:
        h1 = ptg.histogram2d(x, y, weights=w, range=((0, 1), (0, 1)), bins=(300, 200), shape=2)[0]
        h2 = ptg.histogram2d(x, y, weights=w, range=((0, 1), (0, 1)), bins=(300, 200), shape=2)[0]
    finally:
        ptg.method = None
    assert np.array_equal(h1, h2)
    assert abs(h1.base.sum() - w.sum()) <= 1e-12 * w.sum()


def test_fixed_carry_representation_gives_the_same_bits():
    """Beyond 2^33 contributions per cell the accumulators switch to the 128-bit carry form; forced here through the
    PTG_FIXED_CARRY test hook in a child process: both forms must produce the very same grid."""
    import hashlib
    import os
    import subprocess
    import sys
    code = r'''
import hashlib, sys, torch
sys.path.insert(0, %r)
from postpic_b200 import device
g = torch.Generator(device='cuda').manual_seed(99)
n = 3_000_000
x = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
y = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
w = torch.rand(n, dtype=torch.float64, device='cuda', generator=g) - 0.4
out = []
for (cx, cy) in ((x, y), (0.45 + 0.1 * x, 0.45 + 0.1 * y)):
    grid, _ = device.deposit([cx, cy], w, range=((0, 1), (0, 1)), bins=(512, 384), shape=2, method='fixed')
    out.append(hashlib.sha256(grid.cpu().numpy().tobytes()).hexdigest())
print("HASH", *out)
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for tag, env in (("split", {}), ("carry", {"PTG_FIXED_CARRY": "1"})):
        e = dict(os.environ)
        e.update(env)
        p = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        res[tag] = [l for l in p.stdout.splitlines() if l.startswith("HASH")][0]
    assert res["split"] == res["carry"]


def test_fixed_slab_passes_on_a_grid_larger_than_l2():
    """Accumulators beyond 96 MB (here 192 x 192 x 176 cells x 16 B = 104 MB) are deposited in slab passes over axis 0
    (ptg_deposit_fixed.cu).  Integer sums: the grid must carry the same bits whatever the slab size (PTG_FIXED_SLAB_MB
    is read once per process, hence the child processes) and whatever the particle order, and agree with the oracle."""
    import os
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from postpic_b200 import device
rng = np.random.default_rng(4242)
n = 5_000_000
c = [rng.uniform(-0.03, 1.03, n) for _ in range(3)]
w = rng.uniform(-0.4, 1.0, n)
kw = dict(range=[(0, 1)] * 3, bins=(190, 190, 174), shape=1, method='fixed')
dc = [torch.from_numpy(a).cuda() for a in c]
dw = torch.from_numpy(w).cuda()
g1, _ = device.deposit(dc, dw, **kw)
perm = torch.randperm(n, device='cuda')
g2, _ = device.deposit([a[perm] for a in dc], dw[perm], **kw)
assert torch.equal(g1, g2)
np.save(sys.argv[1], g1.cpu().numpy())
''' % root
    grids = {}
    with tempfile.TemporaryDirectory() as tmp:
        for mb in ("16", "72", "4096"):                       # 7 passes, 2 passes, a single pass
            e = dict(os.environ)
            e["PTG_FIXED_SLAB_MB"] = mb
            out = os.path.join(tmp, "g%s.npy" % mb)
            p = subprocess.run([sys.executable, "-c", code, out], env=e, capture_output=True, text=True, timeout=300)
            assert p.returncode == 0, p.stderr[-2000:]
            grids[mb] = np.load(out)
    assert np.array_equal(grids["16"], grids["72"]) and np.array_equal(grids["16"], grids["4096"])
    from oracle import ptg_oracle
    rng = np.random.default_rng(4242)
    n = 5_000_000
    c = [rng.uniform(-0.03, 1.03, n) for _ in range(3)]
    w = rng.uniform(-0.4, 1.0, n)
    ho, _ = cases.call(ptg_oracle, c, w, 1, ranges=[(0, 1)] * 3, bins=(190, 190, 174))
    b = cases.padded(ho)
    a = grids["72"].reshape(b.shape)
    assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b))
