"""CUDA vs oracle on the geometry of the BASELINE.json configurations.

cfg2 (1024^2 TSC, x in U[0,1), px in U[-1,1)), cfg3 (256^3 CIC), cfg4 (2048^2 spline3 on a 40 um box,
weights = weight * Ekin_MeV, precomputed and fused into the kernel) and cfg5 (clustered TSC: 90 % of
the particles in 1 % of the cells) at 1e7 particles, in random, x-sorted and cell-sorted order, through
AUTO and through every method AUTO can pick for such input (plus GLOBAL and the deterministic SORTED).
Per bin |gpu - oracle| <= 1e-12 max|oracle|, in total <= 1e-12 sum|w| (the north-star bar); shape 0 with
integer weights bit-exact at the same sizes; the reference's own invariant (mass incl. ghost cells,
/root/reference/test/test_particlestogrid.py:54) on top.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = int(float(os.environ.get("PTG_TEST_N", "1e7")))
ME, C0, QE = 9.1093837015e-31, 299792458.0, 1.602176634e-19

CONFIGS = {
    # name: (ndim, shape, bins, ranges)
    "cfg2": (2, 2, (1024, 1024), ((0.0, 1.0), (-1.0, 1.0))),
    "cfg3": (3, 1, (256, 256, 256), ((0.0, 1.0),) * 3),
    "cfg4": (2, 3, (2048, 2048), ((0.0, 40e-6), (0.0, 40e-6))),
    "cfg5": (2, 2, (1024, 1024), ((0.0, 1.0), (0.0, 1.0))),
}
ORDERS = ("random", "xsorted", "cellsorted")
METHODS = (None, "rows", "tiled", "runs", "runs_wide", "global", "sorted", "fixed")

_cache = {}


def _ekin_mev(px, py, pz, mass):
    # scalarproperties.py:183-205 in the reference's operation order
    np2 = (px ** 2 + py ** 2 + pz ** 2) / (mass * C0) ** 2
    gm1 = np2 / (np.sqrt(1 + np2) + 1)
    return gm1 * mass * C0 ** 2 / QE / 1e6


def cloud(name, order, n=None, seed=None):
    """Seeded host cloud of the configuration (numpy, so that the oracle sees the very same doubles)."""
    n = N if n is None else n
    key = (name, order, n, seed)
    if key in _cache:
        return _cache[key]
    _cache.clear()                      # one cloud at a time: 1e7 x 7 arrays is enough host memory
    nd, shape, bins, ranges = CONFIGS[name]
    rng = np.random.default_rng({"cfg2": 2, "cfg3": 3, "cfg4": 4, "cfg5": 5}[name] if seed is None else seed)
    coords = []
    pick = rng.random(n) < 0.9 if name == "cfg5" else None      # 90 % of the particles in 1 % of the cells
    for a in range(nd):
        lo, hi = ranges[a]
        u = rng.random(n)
        if pick is not None:
            u = np.where(pick, 0.45 + 0.1 * u, u)
        coords.append(lo + (hi - lo) * u)
    extra = {}
    if name == "cfg4":
        mom = [0.1 * ME * C0 * rng.standard_normal(n) for _ in range(3)]
        weight = np.full(n, 3.0e7)                                  # macro-particle weight, constant
        w = weight * _ekin_mev(mom[0], mom[1], mom[2], ME)
        extra = dict(mom=mom, weight=weight)
    else:
        w = rng.random(n)
    if order != "random":
        if order == "xsorted":
            idx = np.argsort(coords[0], kind="stable")
        else:
            k = np.zeros(n, dtype=np.int64)
            for a in range(nd):
                lo, hi = ranges[a]
                k = k * bins[a] + np.clip(np.floor((coords[a] - lo) / (hi - lo) * bins[a]), 0, bins[a] - 1).astype(np.int64)
            idx = np.argsort(k, kind="stable")
        coords = [np.ascontiguousarray(c[idx]) for c in coords]
        w = np.ascontiguousarray(w[idx])
        if extra:
            extra = dict(mom=[np.ascontiguousarray(m[idx]) for m in extra["mom"]], weight=extra["weight"])
    _cache[key] = (coords, w, extra)
    return _cache[key]


def oracle_padded(oracle, name, coords, w, shape=None):
    nd, s, bins, ranges = CONFIGS[name]
    s = s if shape is None else shape
    if nd == 2:
        h = oracle.histogram2d(coords[0], coords[1], weights=w, range=ranges, bins=bins, shape=s)[0]
    else:
        h = oracle.histogram3d(coords[0], coords[1], coords[2], weights=w, range=ranges, bins=bins, shape=s)[0]
    return h.base if h.base is not None else h


def check(got, ref, wsum_abs, tag):
    assert got.shape == ref.shape, tag
    scale = float(np.max(np.abs(ref)))
    per_bin = float(np.max(np.abs(got - ref)))
    assert per_bin <= 1e-12 * scale, "%s: per-bin %.3e > 1e-12 * %.3e" % (tag, per_bin, scale)
    assert abs(got.sum() - ref.sum()) <= 1e-12 * wsum_abs, tag
    return per_bin / scale


@pytest.mark.parametrize("order", ORDERS)
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_baseline_geometry_matches_oracle(oracle, name, order):
    import torch
    from postpic_b200 import device
    nd, shape, bins, ranges = CONFIGS[name]
    coords, w, extra = cloud(name, order)
    ref = oracle_padded(oracle, name, coords, w)
    wsum = float(np.abs(w).sum())
    # the reference's own invariant: everything lies inside the range, so the padded array holds all the weight
    assert abs(ref.sum() - w.sum()) <= 1e-12 * wsum
    dc = [torch.from_numpy(c).cuda() for c in coords]
    dw = torch.from_numpy(w).cuda()
    kw = dict(range=ranges, bins=bins, shape=shape)
    worst = {}
    for m in METHODS:
        if m == "runs_wide" and ((shape + 1) + 1) ** nd > 27:
            continue
        grid, plan = device.deposit(dc, dw, method=m, **kw)
        worst[m] = check(grid.cpu().numpy(), ref, wsum, "%s/%s/%s" % (name, order, m))
    if name == "cfg4":
        # the same field with the weight expression evaluated inside the kernel (createField(weights='Ekin_MeV'))
        dm = [torch.from_numpy(m).cuda() for m in extra["mom"]]
        dwt = torch.from_numpy(extra["weight"]).cuda()
        for m in (None, "rows", "tiled", "runs"):
            grid, plan = device.deposit(dc, dwt, method=m, fused=dict(kind="Ekin_MeV", px=dm[0], py=dm[1], pz=dm[2],
                                                                        mass=ME, c=C0, elementary_charge=QE), **kw)
            worst["fused/%s" % m] = check(grid.cpu().numpy(), ref, wsum, "%s/%s/fused/%s" % (name, order, m))
    print("%s %s: worst per-bin / max = %s" % (name, order, {str(k): "%.1e" % v for k, v in worst.items()}))


@pytest.mark.parametrize("order", ORDERS)
@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg5"])
def test_baseline_geometry_shape0_integer_weights_bit_exact(oracle, name, order):
    import torch
    from postpic_b200 import device
    nd, _, bins, ranges = CONFIGS[name]
    coords, _, _ = cloud(name, order)
    w = np.random.default_rng(11).integers(1, 9, coords[0].size).astype(np.float64)
    ref = oracle_padded(oracle, name, coords, w, shape=0)
    dc = [torch.from_numpy(c).cuda() for c in coords]
    dw = torch.from_numpy(w).cuda()
    for m in (None, "global", "tiled", "runs", "runs_wide", "sorted", "fixed"):
        grid, _ = device.deposit(dc, dw, range=ranges, bins=bins, shape=0, method=m)
        assert np.array_equal(grid.cpu().numpy(), ref), "%s/%s/%s" % (name, order, m)
    # and plain counting (weights=None)
    ref = oracle_padded(oracle, name, coords, None, shape=0)
    for m in (None, "tiled", "runs"):
        grid, _ = device.deposit(dc, None, range=ranges, bins=bins, shape=0, method=m)
        assert np.array_equal(grid.cpu().numpy(), ref), "%s/%s/%s/count" % (name, order, m)


def test_numpy_drop_in_on_cfg2_geometry(oracle):
    """The host-buffer entry (ptg_histogram_host_f64: chunked H2D pipeline, several chunks) on the headline geometry."""
    from postpic_b200.particles import _particlestogrid as ptg
    coords, w, _ = cloud("cfg2", "random")
    nd, shape, bins, ranges = CONFIGS["cfg2"]
    ref = oracle_padded(oracle, "cfg2", coords, w)
    ptg.method = None
    h, ex, ey = ptg.histogram2d(coords[0], coords[1], weights=w, range=ranges, bins=bins, shape=shape)
    check(h.base, ref, float(np.abs(w).sum()), "cfg2 numpy drop-in")
    eo = oracle.histogram2d(coords[0][:10], coords[1][:10], weights=w[:10], range=ranges, bins=bins, shape=shape)
    assert np.array_equal(ex, eo[1]) and np.array_equal(ey, eo[2])
