"""Parity of the CUDA path with the CPU oracle, through the Python drop-in and thus the
C ABI (ptg_histogram_host_f64 for numpy inputs, ptg_deposit_f64 for CUDA tensors).

Tolerances (BASELINE.json north_star):
  * shape 0 with integer-valued (or no) weights: BIT-EXACT (np.array_equal);
  * shapes 1-3: |sum(gpu) - sum(ref)| <= 1e-12 * sum|w_accepted|   (the stated bar), and
    in addition per bin |gpu - ref| <= 1e-12 * max|ref| (every per-particle contribution
    is computed with the reference's exact operation order, only the summation order
    differs, so the observed differences are ~1e-16 relative).
"""
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "ptg_golden.npz"))
ALL = [(nd, s, w) for nd in (1, 2, 3) for s in (0, 1, 2, 3) for w in (False, True)]
METHODS = [None, 'global', 'rows', 'tiled', 'tiled_exact', 'sorted', 'fixed', 'runs', 'runs_wide']


def _skip_unavailable(method, nd, shape):
    # the widened-window runs kernel exists only where (S+2)^ndim <= 27 accumulators fit
    if method == 'runs_wide' and ((1 if shape == 0 else shape + 1) + 1) ** nd > 27:
        pytest.skip("no runs_wide instance for ndim=%d shape=%d" % (nd, shape))


@pytest.fixture
def ptg():
    """The drop-in module with its default method; tests that force a method set `ptg.method` themselves
    and never leak it into the next test."""
    from postpic_b200.particles import _particlestogrid
    _particlestogrid.method = None
    yield _particlestogrid
    _particlestogrid.method = None


def assert_parity(h_gpu, h_ref, wsum_abs, exact=False):
    a, b = cases.padded(h_gpu), cases.padded(h_ref)
    assert a.shape == b.shape
    if exact:
        assert np.array_equal(a, b)
        return
    assert abs(a.sum() - b.sum()) <= 1e-12 * max(wsum_abs, 1e-300)
    scale = np.max(np.abs(b)) if b.size else 0.0
    assert np.max(np.abs(a - b), initial=0.0) <= 1e-12 * scale + 1e-300


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("nd,shape,weighted", ALL)
def test_matches_oracle_seeded(ptg, oracle, nd, shape, weighted, method):
    _skip_unavailable(method, nd, shape)
    ptg.method = method
    coords, w = cases.cloud(4242 + 7 * nd + shape, 200000, nd, weighted, spill=0.15)
    hg, eg = cases.call(ptg, coords, w, shape)
    ho, eo = cases.call(oracle, coords, w, shape)
    _, asum = oracle.accepted_weight(coords, w, cases.RANGES[:nd], cases.BINS[:nd], shape)
    assert_parity(hg, ho, asum, exact=(shape == 0 and not weighted))
    assert hg.shape == ho.shape and hg.strides == ho.strides and hg.dtype == np.float64
    for a, b in zip(eg, eo):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("nd", [1, 2, 3])
def test_shape0_integer_weights_bit_exact(ptg, oracle, nd):
    coords, w = cases.cloud(99 + nd, 300000, nd, True, integer_weights=True)
    hg, _ = cases.call(ptg, coords, w, 0)
    ho, _ = cases.call(oracle, coords, w, 0)
    assert np.array_equal(hg, ho)


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("nd,shape,weighted", ALL)
def test_matches_golden(ptg, nd, shape, weighted, method):
    _skip_unavailable(method, nd, shape)
    ptg.method = method
    key = "d%d_s%d_w%d" % (nd, shape, int(weighted))
    coords, w = cases.cloud(int(GOLDEN[key + "_seed"]), 4000, nd, weighted)
    h, edges = cases.call(ptg, coords, w, shape)
    ref = GOLDEN[key + "_padded"]
    got = cases.padded(h)
    if shape == 0 and not weighted:
        assert np.array_equal(got, ref)
    else:
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
        assert abs(got.sum() - ref.sum()) <= 1e-12 * np.abs(ref).sum()
    for a, e in enumerate(edges):
        assert np.array_equal(e, GOLDEN[key + "_edges%d" % a])


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("shape", [0, 1, 2, 3])
def test_special_values_golden(ptg, shape, method):
    if method == 'runs_wide' and shape >= 2:
        pytest.skip("no 3-D runs_wide instance for this shape")
    ptg.method = method
    sp = GOLDEN["special_values"]
    good = np.full_like(sp, 0.37)
    ysafe = GOLDEN["special_ysafe"]

    def close(h, key):
        ref = GOLDEN[key]
        assert np.max(np.abs(cases.padded(h) - ref)) <= 1e-12 * max(np.max(np.abs(ref)), 1e-300)
    h, _ = ptg.histogram(sp, range=(0, 1), bins=10, shape=shape)
    close(h, "special_d1_s%d" % shape)
    h, _, _ = ptg.histogram2d(sp, good, range=((0, 1), (0, 1)), bins=(10, 7), shape=shape)
    close(h, "special_d2x_s%d" % shape)
    h, _, _ = ptg.histogram2d(good, ysafe, range=((0, 1), (0, 1)), bins=(7, 10), shape=shape)
    close(h, "special_d2y_s%d" % shape)
    h, _, _, _ = ptg.histogram3d(good, good, ysafe, range=((0, 1),) * 3, bins=(5, 6, 10), shape=shape)
    close(h, "special_d3z_s%d" % shape)


@pytest.mark.parametrize("shape", [0, 1, 2, 3])
def test_degenerate_range(ptg, shape):
    h, e = ptg.histogram(GOLDEN["degenerate_data"], range=(2.0, 2.0), bins=4, shape=shape)
    ref = GOLDEN["degenerate_s%d" % shape]
    assert np.allclose(cases.padded(h), ref, rtol=1e-12, atol=0)
    assert np.array_equal(e, GOLDEN["degenerate_edges"])


@pytest.mark.parametrize("shape", [0, 1, 2, 3])
def test_known_answers(ptg, shape):
    h, e = ptg.histogram(cases.KNOWN_1D_POS, range=(0, 25), bins=25, shape=shape)
    expect = np.zeros(25)
    for k, v in cases.KNOWN_1D[shape].items():
        expect[k] = v
    np.testing.assert_allclose(h, expect, rtol=1e-15, atol=1e-17)
    h, _, _ = ptg.histogram2d(np.array([2.25]), np.array([3.6]), weights=np.array([2.0]),
                              range=((0, 5), (0, 6)), bins=(5, 6), shape=2)
    expect = np.zeros((5, 6))
    expect[1:4, 2:5] = cases.KNOWN_2D_BLOCK
    np.testing.assert_allclose(h, expect, rtol=1e-13, atol=1e-16)


def test_single_particle_contributions_bit_exact(ptg, oracle):
    # one particle => no summation-order freedom: the GPU stencil must equal the oracle's bit for bit
    rng = np.random.default_rng(3)
    for shape in (1, 2, 3):
        for _ in range(25):
            x, y, z, w = rng.uniform(-0.05, 1.05, 4)
            args = dict(range=((0, 1),) * 3, bins=(6, 5, 4), shape=shape, weights=np.array([w]))
            hg = ptg.histogram3d(np.array([x]), np.array([y]), np.array([z]), **args)[0]
            ho = oracle.histogram3d(np.array([x]), np.array([y]), np.array([z]), **args)[0]
            assert np.array_equal(cases.padded(hg), cases.padded(ho))


def test_undefined_reference_inputs_are_dropped(ptg):
    bad = np.array([1e300, np.inf, 3e9])
    good = np.full(3, 0.5)
    for shape in (1, 2, 3):
        h, _, _ = ptg.histogram2d(good, bad, range=((0, 1), (0, 1)), bins=(4, 4), shape=shape)
        assert cases.padded(h).sum() == 0.0
        h, _, _, _ = ptg.histogram3d(good, good, bad, range=((0, 1),) * 3, bins=(4, 4, 4), shape=shape)
        assert cases.padded(h).sum() == 0.0


def test_empty_and_tiny_inputs(ptg, oracle):
    e = np.empty(0)
    for shape in (0, 1, 2, 3):
        h, edges = ptg.histogram(e, range=(0, 1), bins=7, shape=shape)
        assert h.shape == (7,) and not h.any() and np.array_equal(edges, np.linspace(0, 1, 8))
        h, _, _ = ptg.histogram2d(e, e, range=((0, 1), (0, 1)), bins=(3, 4), shape=shape, weights=e)
        assert h.shape == (3, 4) and not h.any()
        one = np.array([0.4])
        hg, _, _, _ = ptg.histogram3d(one, one, one, range=((0, 1),) * 3, bins=(3, 4, 5), shape=shape)
        ho, _, _, _ = oracle.histogram3d(one, one, one, range=((0, 1),) * 3, bins=(3, 4, 5), shape=shape)
        assert np.array_equal(hg, ho)


def test_ragged_sizes_around_warp_and_chunk_boundaries(ptg, oracle):
    rng = np.random.default_rng(8)
    for n in (1, 31, 32, 33, 255, 256, 257, 1023, 1025, 4097):
        x, y = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
        w = rng.uniform(0, 1, n)
        for shape in (1, 2, 3):
            hg = ptg.histogram2d(x, y, weights=w, range=((0, 1), (0, 1)), bins=(8, 9), shape=shape)[0]
            ho = oracle.histogram2d(x, y, weights=w, range=((0, 1), (0, 1)), bins=(8, 9), shape=shape)[0]
            assert_parity(hg, ho, w.sum())


def test_multi_chunk_host_pipeline(ptg, oracle):
    # > 2 pipeline chunks (chunk = 2^22 particles) with a ragged tail, through the default method and
    # through the run-length kernels (unordered input: every particle closes a run)
    n = (1 << 23) + 12345
    rng = np.random.default_rng(12)
    x, y, w = rng.uniform(-0.01, 1.01, n), rng.uniform(-0.01, 1.01, n), rng.uniform(0, 1, n)
    ho = oracle.histogram2d(x, y, weights=w, range=((0, 1), (0, 1)), bins=(128, 96), shape=2)[0]
    _, asum = oracle.accepted_weight([x, y], w, [(0, 1), (0, 1)], [128, 96], 2)
    for method in (None, 'runs', 'runs_wide'):
        ptg.method = method
        hg = ptg.histogram2d(x, y, weights=w, range=((0, 1), (0, 1)), bins=(128, 96), shape=2)[0]
        assert_parity(hg, ho, asum)
    ptg.method = None
    # shape 0, unweighted, same size: bit exact
    hg = ptg.histogram2d(x, y, range=((0, 1), (0, 1)), bins=(128, 96), shape=0)[0]
    ho = oracle.histogram2d(x, y, range=((0, 1), (0, 1)), bins=(128, 96), shape=0)[0]
    assert np.array_equal(hg, ho)


def test_strided_and_reversed_inputs(ptg, oracle):
    base = np.random.default_rng(5).uniform(0, 1, (5000, 3))
    cols = [base[:, i] for i in range(3)]
    rev = [c[::-1] for c in cols]
    for shape in (0, 2):
        a, _ = cases.call(ptg, cols, None, shape, ranges=[(0, 1)] * 3, bins=[9, 8, 7])
        b, _ = cases.call(oracle, cols, None, shape, ranges=[(0, 1)] * 3, bins=[9, 8, 7])
        assert_parity(a, b, 5000.0, exact=(shape == 0))
        c, _ = cases.call(ptg, rev, None, shape, ranges=[(0, 1)] * 3, bins=[9, 8, 7])
        assert_parity(c, b, 5000.0, exact=(shape == 0))


def test_argument_errors(ptg):
    x = np.linspace(0, 1, 10)
    with pytest.raises(ValueError):
        ptg.histogram2d(x, x[:5], range=((0, 1), (0, 1)), bins=(4, 4))
    with pytest.raises(ValueError):
        ptg.histogram3d(x, x, x[:5], range=((0, 1),) * 3, bins=(4, 4, 4))
    with pytest.raises(ValueError):
        ptg.histogram(x.astype(np.float32), range=(0, 1), bins=4)
    with pytest.raises(ValueError):
        ptg.histogram(x.reshape(2, 5), range=(0, 1), bins=4)
    with pytest.raises(TypeError):
        ptg.histogram(list(x), range=(0, 1), bins=4)
    with pytest.raises(TypeError):
        ptg.histogram(x, range=None, bins=4)
    with pytest.raises(ValueError):
        ptg.histogram(x, range=(0, 1), bins=4, shape=5)
    with pytest.raises(ValueError):
        ptg.histogram(x, range=(0, 1), bins=4, weights=np.ones(3), shape=1)


def test_shape_spellings_and_view_contract(ptg):
    x = np.random.default_rng(1).uniform(0, 1, 1000)
    ref = ptg.histogram(x, range=(0, 1), bins=10, shape=2)[0]
    # (atomic accumulation: two runs agree to rounding, not bit for bit)
    for spelling in ('triangle', 2.0, np.int64(2)):
        assert np.allclose(ptg.histogram(x, range=(0, 1), bins=10, shape=spelling)[0], ref, rtol=1e-13, atol=0)
    assert np.allclose(ptg.histogram(x, range=(0, 1), bins=10, shape=True)[0],
                       ptg.histogram(x, range=(0, 1), bins=10, shape='tophat')[0], rtol=1e-13, atol=0)
    h2 = ptg.histogram2d(x, x, range=((0, 1), (0, 1)), bins=(10, 10), shape=2)[0]
    assert h2.base is not None and h2.base.shape == (14, 14) and h2.strides == (112, 8)
    assert not h2.flags['OWNDATA']
    h2 /= 3.0           # the caller's in-place scaling must work (particles.py:1102)


def test_device_tensor_inputs_match_host_inputs(ptg):
    import torch
    rng = np.random.default_rng(21)
    n = 100000
    x, y, w = rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    tx, ty, tw = (torch.from_numpy(a).cuda() for a in (x, y, w))
    for shape in (0, 1, 2, 3):
        a = ptg.histogram2d(x, y, weights=w, range=((0, 1), (0, 1)), bins=(64, 48), shape=shape)[0]
        b = ptg.histogram2d(tx, ty, weights=tw, range=((0, 1), (0, 1)), bins=(64, 48), shape=shape)[0]
        assert_parity(b, a, w.sum())
    # strided device views of an (N, 2) tensor
    t2 = torch.stack([tx, ty], dim=1)
    c = ptg.histogram2d(t2[:, 0], t2[:, 1], weights=tw, range=((0, 1), (0, 1)), bins=(64, 48), shape=2)[0]
    a = ptg.histogram2d(x, y, weights=w, range=((0, 1), (0, 1)), bins=(64, 48), shape=2)[0]
    assert_parity(c, a, w.sum())


def test_div6_bit_exact():
    import ctypes
    import torch
    from postpic_b200 import _lib
    rng = np.random.default_rng(6)
    vals = np.concatenate([rng.uniform(0, 1, 2_000_000), rng.uniform(0, 1, 2_000_000) ** 3,
                           rng.uniform(-4, 4, 1_000_000), 10.0 ** rng.uniform(-200, 200, 1_000_000)])
    tin = torch.from_numpy(vals).cuda()
    tout = torch.empty_like(tin)
    _lib.check(_lib.load().ptg_selftest_div6(tin.numel(), tin.data_ptr(), tout.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(tout.cpu().numpy(), vals / 6.0)


def test_minmax_device_and_host():
    import ctypes
    import torch
    from postpic_b200 import _lib, device
    rng = np.random.default_rng(9)
    a = rng.normal(0, 5, 3_000_001)
    assert device.minmax(torch.from_numpy(a).cuda()) == (a.min(), a.max())
    assert device.minmax(torch.from_numpy(a).cuda()[::3]) == (a[::3].min(), a[::3].max())
    out = np.zeros(2)
    big = rng.normal(0, 5, (1 << 23) + 77)
    _lib.check(_lib.load().ptg_minmax_host_f64(big.size, big.ctypes.data, 1, out.ctypes.data, -1))
    assert out[0] == big.min() and out[1] == big.max()
    a[12345] = np.nan
    lo, hi = device.minmax(torch.from_numpy(a).cuda())
    assert np.isnan(lo) and np.isnan(hi)


# ---- size-independent properties at a size the oracle would take minutes for ----------
@pytest.mark.parametrize("shape", [0, 2, 3])
def test_large_properties_on_device(shape):
    import torch
    from postpic_b200 import device
    n = 50_000_000
    g = torch.Generator(device='cuda').manual_seed(1234 + shape)
    x = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    y = torch.rand(n, dtype=torch.float64, device='cuda', generator=g) * 2 - 1
    w = torch.randint(1, 5, (n,), device='cuda', generator=g).to(torch.float64)
    kw = dict(range=((0, 1), (-1, 1)), bins=(1024, 1024), shape=shape)
    grid, plan = device.deposit([x, y], w, **kw)
    total = float(w.sum())
    # mass conservation including ghost cells (test/test_particlestogrid.py:54): every particle is inside
    if shape == 0:
        inside = (x > 0) & (y > -1)
        assert float(grid.sum()) == float(w[inside].sum())
    else:
        assert abs(float(grid.sum()) - total) <= 1e-12 * total
    # linearity: depositing the two halves separately and adding equals the whole
    h = n // 2
    g1, _ = device.deposit([x[:h], y[:h]], w[:h], **kw)
    g2, _ = device.deposit([x[h:], y[h:]], w[h:], **kw)
    if shape == 0:
        assert torch.equal(g1 + g2, grid)          # integer weights: exact under any order
    else:
        assert float((g1 + g2 - grid).abs().max()) <= 1e-12 * float(grid.abs().max())
    # permutation invariance (the reference is order-sensitive only at the 1e-15 level)
    perm = torch.randperm(n, device='cuda', generator=g)
    gp, _ = device.deposit([x[perm], y[perm]], w[perm], **kw)
    if shape == 0:
        assert torch.equal(gp, grid)
    else:
        assert float((gp - grid).abs().max()) <= 1e-12 * float(grid.abs().max())
    # methods agree
    if shape:
        ga, _ = device.deposit([x, y], w, method='global', **kw)
        for m in ('rows', 'tiled', 'sorted', 'runs', 'runs_wide'):
            gb, _ = device.deposit([x, y], w, method=m, **kw)
            assert float((ga - gb).abs().max()) <= 1e-12 * float(grid.abs().max())
        # x-sorted and cell-sorted orders exercise the tile window moves and the warp aggregation
        order = torch.argsort(x)
        gs, _ = device.deposit([x[order], y[order]], w[order], method='tiled', **kw)
        assert float((gs - grid).abs().max()) <= 1e-12 * float(grid.abs().max())
        cell = (x * 1024).floor() * 1024 + ((y + 1) * 512).floor()
        order = torch.argsort(cell)
        gc, _ = device.deposit([x[order], y[order]], w[order], method='tiled', **kw)
        assert float((gc - grid).abs().max()) <= 1e-12 * float(grid.abs().max())
        for m in (None, 'runs', 'runs_wide'):
            gc, _ = device.deposit([x[order], y[order]], w[order], method=m, **kw)
            assert float((gc - grid).abs().max()) <= 1e-12 * float(grid.abs().max())
        # unaligned / strided views take the scalar-load variant of the runs kernel
        gc, _ = device.deposit([x[order][1:], y[order][1:]], w[order][1:], method='runs', **kw)
        g1, _ = device.deposit([x[order][1:], y[order][1:]], w[order][1:], method='global', **kw)
        assert float((gc - g1).abs().max()) <= 1e-12 * float(grid.abs().max())


@pytest.mark.parametrize("nd,shape", [(1, 1), (2, 0), (2, 2), (2, 3), (3, 1), (3, 3)])
def test_sorted_method_is_bit_reproducible(nd, shape):
    """method 'sorted' (sort by cell + segmented reduce, no atomics): repeated runs are bitwise
    identical, also on a clustered cloud where the atomic paths contend most."""
    import torch
    from postpic_b200 import device
    n = 3_000_000
    g = torch.Generator(device='cuda').manual_seed(5 + nd + shape)
    coords = []
    for a in range(nd):
        u = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
        pick = torch.rand(n, device='cuda', generator=g) < 0.9
        coords.append(torch.where(pick, 0.45 + 0.1 * u, u * 1.1 - 0.05))
    w = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    bins = [(4096,), (512, 384), (96, 80, 64)][nd - 1]
    kw = dict(range=[(0, 1)] * nd, bins=bins, shape=shape)
    g1, plan = device.deposit(coords, w, method='sorted', **kw)
    g2, _ = device.deposit(coords, w, method='sorted', **kw)
    g3, _ = device.deposit(coords, w, method='sorted', **kw)
    assert torch.equal(g1, g2) and torch.equal(g1, g3)
    ga, _ = device.deposit(coords, w, method='global', **kw)
    assert float((g1 - ga).abs().max()) <= 1e-12 * float(ga.abs().max())


def test_slab_passes_on_a_grid_larger_than_l2():
    """3-D grid of 242^3 doubles = 113 MB: methods rows / global deposit it in slab passes over
    axis 0; the un-slabbed runs kernel is the cross-check."""
    import torch
    from postpic_b200 import device
    n = 6_000_000
    g = torch.Generator(device='cuda').manual_seed(77)
    c = [torch.rand(n, dtype=torch.float64, device='cuda', generator=g) * 1.02 - 0.01 for _ in range(3)]
    w = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    kw = dict(range=[(0, 1)] * 3, bins=(240, 240, 240), shape=1)
    ref, plan = device.deposit(c, w, method='runs', **kw)
    for m in ('rows', 'global', None):
        got, _ = device.deposit(c, w, method=m, **kw)
        assert float((got - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
        assert abs(float(got.sum()) - float(ref.sum())) <= 1e-12 * float(w.sum())


@pytest.mark.parametrize("nd,shape", [(1, 1), (1, 3), (2, 1), (2, 3), (3, 1), (2, 2)])
def test_cell_ordered_input_odd_shapes(nd, shape):
    """Cell-ordered particles (sorted by floor cell): for odd shapes the stencil base cell flips
    between f and f+1 inside a cell, which the widened-window runs kernel absorbs.  AUTO, runs
    and runs_wide must all agree with the atomic reference path, incl. non-finite weights."""
    import torch
    from postpic_b200 import device
    n = 4_000_000
    g = torch.Generator(device='cuda').manual_seed(900 + 10 * nd + shape)
    bins = [(4096,), (300, 200), (48, 40, 36)][nd - 1]
    coords = [torch.rand(n, dtype=torch.float64, device='cuda', generator=g) * 1.04 - 0.02 for _ in range(nd)]
    key = torch.zeros(n, dtype=torch.int64, device='cuda')
    for a in range(nd):
        key = key * bins[a] + (coords[a] * bins[a]).floor().clamp(0, bins[a] - 1).to(torch.int64)
    order = torch.argsort(key)
    coords = [c[order].contiguous() for c in coords]
    w = torch.rand(n, dtype=torch.float64, device='cuda', generator=g) - 0.3
    kw = dict(range=[(0, 1)] * nd, bins=bins, shape=shape)
    ref, _ = device.deposit(coords, w, method='global', **kw)
    for m in (None, 'runs', 'runs_wide', 'tiled'):
        got, _ = device.deposit(coords, w, method=m, **kw)
        assert float((got - ref).abs().max()) <= 1e-12 * float(ref.abs().max()), m
    # a few inf / NaN weights must poison exactly the cells the reference poisons
    w2 = w.clone()
    w2[12345] = float('inf')
    w2[2_000_001] = float('nan')
    ref, _ = device.deposit(coords, w2, method='global', **kw)
    for m in (None, 'runs_wide'):
        got, _ = device.deposit(coords, w2, method=m, **kw)
        assert torch.equal(torch.isfinite(got), torch.isfinite(ref)), m
        fin = torch.isfinite(ref)
        assert float((got[fin] - ref[fin]).abs().max()) <= 1e-12 * float(ref[fin].abs().max()), m


def test_auto_probe_verdicts():
    """AUTO's device-side probe picks the method the input order calls for."""
    import torch
    from postpic_b200 import device, _lib
    n = 8_000_000
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    y = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    kw = dict(range=[(0, 1), (0, 1)], bins=(1024, 1024))
    device.deposit([x, y], None, shape=2, **kw)
    assert _lib.last_probe()['mode'] == 'rows'                       # i.i.d. over a grid >> tile
    o = torch.argsort(x)
    device.deposit([x[o], y[o]], None, shape=2, **kw)
    assert _lib.last_probe()['mode'] == 'tiled'                      # x-ordered: a band of whole rows in the fixed-point tile
    device.deposit([x[o], y[o]], None, shape=2, method='auto_exact', **kw)
    assert _lib.last_probe()['mode'] == 'rows'                       # ... but no win with the fp64 CAS tile
    cell = (x * 1024).floor() * 1024 + (y * 1024).floor()
    o = torch.argsort(cell)
    xs, ys = x[o].contiguous(), y[o].contiguous()
    device.deposit([xs, ys], None, shape=2, **kw)
    assert _lib.last_probe()['mode'] == 'runs'                       # cell-ordered, even shape
    device.deposit([xs, ys], None, shape=1, **kw)
    assert _lib.last_probe()['mode'] == 'runs_wide'                  # cell-ordered, odd shape
    xb = 0.45 + 0.1 * x
    yb = 0.45 + 0.1 * y
    device.deposit([xb, yb], None, shape=2, **kw)
    p = _lib.last_probe()
    assert p['mode'] == 'tiled' and p['ext'][0] * p['ext'][1] <= 25600     # a bunch that fits the tile
    device.deposit([x], None, range=(0, 1), bins=1000, shape=1)
    p = _lib.last_probe()
    assert p['mode'] == 'tiled' and p['ext'][0] == 1002              # small 1-D histogram: whole grid in the tile


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("nd,shape", [(nd, s) for nd in (1, 2, 3) for s in (0, 1, 2, 3)])
def test_no_write_outside_the_padded_grid(nd, shape, method):
    """Guard zones on both sides of the grid must stay untouched (compute-sanitizer is not available on
    this pool, so out-of-bounds reductions are hunted with canaries): particles overshoot the range on
    every side, sit exactly on the edges, and include NaN / inf coordinates."""
    import torch
    from postpic_b200 import device
    _skip_unavailable(method, nd, shape)
    bins = [(61,), (23, 17), (9, 8, 7)][nd - 1]
    plan = device.Plan([(0.0, 1.0)] * nd, list(bins), shape)
    guard = 8192
    buf = torch.zeros(plan.cells + 2 * guard, dtype=torch.float64, device='cuda')
    grid = buf[guard:guard + plan.cells].view(plan.dims)
    g = torch.Generator(device='cuda').manual_seed(17 * nd + shape)
    n = 300000
    coords = [torch.rand(n, dtype=torch.float64, device='cuda', generator=g) * 1.6 - 0.3 for _ in range(nd)]
    special = torch.tensor([0.0, 1.0, -0.0, float('nan'), float('inf'), float('-inf'), 1e300, -1e300, 3e9,
                            1.0 - 1e-16, 1.0 + 1e-12, -1e-12], dtype=torch.float64, device='cuda')
    for c in coords:
        c[:special.numel()] = special
        c[special.numel():2 * special.numel()] = special.flip(0)
    # cell-ordered copy as well, so that the run-length kernels take their aggregated paths
    key = torch.zeros(n, dtype=torch.float64, device='cuda')
    for c in coords:
        key = key * 64 + torch.nan_to_num(c, nan=0.0, posinf=2.0, neginf=-1.0).clamp(-1, 2).mul(20).floor()
    order = torch.argsort(key)
    w = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    for cs, ws in ((coords, w), ([c[order].contiguous() for c in coords], w[order].contiguous()), (coords, None)):
        device.deposit_into(grid, plan, cs, ws, method=method, zero=True)
        torch.cuda.synchronize()
        assert not buf[:guard].any() and not buf[guard + plan.cells:].any()
        assert torch.isfinite(grid).all()


def test_randomised_differential_against_oracle(ptg, oracle):
    """400 random problems (dimension, shape, weights or not, 1..48 bins per axis, forward / reversed /
    tiny / huge ranges, 0..5000 particles that overshoot the range, NaN and inf sprinkled in, every
    deposition method) -- CUDA path vs oracle, padded grids and edges."""
    rng = np.random.default_rng(20240607)
    checked = 0
    for trial in range(400):
        nd = int(rng.integers(1, 4))
        shape = int(rng.integers(0, 4))
        weighted = bool(rng.integers(0, 2))
        method = METHODS[int(rng.integers(0, len(METHODS)))]
        if method == 'runs_wide' and ((1 if shape == 0 else shape + 1) + 1) ** nd > 27:
            method = 'runs'
        n = int(rng.choice([0, 1, 2, 31, 33, 257, 1000, 5000]))
        bins = [int(rng.integers(1, 49)) for _ in range(nd)]
        ranges, coords = [], []
        for a in range(nd):
            kind = rng.integers(0, 5)
            lo = float(rng.normal(0, 10))
            span = float(10.0 ** rng.uniform(-6, 6))
            hi = lo + span
            if kind == 0:
                lo, hi = hi, lo                               # reversed range
            elif kind == 1:
                lo, hi, span = 0.0, 1.0, 1.0
            ranges.append((lo, hi))
            c = rng.uniform(min(lo, hi) - 0.2 * span, max(lo, hi) + 0.2 * span, n)
            if n > 8:
                # raw-double-tested axes (y; z for 3-D shape 1) crash the reference for +inf / huge values
                # (dropped here), so those only go on int-tested axes where the reference is defined
                c[0] = np.nan
                c[1] = -np.inf
                if a == 0 or (a == 2 and shape != 1):
                    c[2] = np.inf
                    c[3] = 1e300
                c[4] = lo
                c[5] = hi
            coords.append(c)
        w = rng.normal(1.0, 2.0, n) if weighted else None
        ptg.method = method
        hg, eg = cases.call(ptg, coords, w, shape, ranges=ranges, bins=bins)
        ho, eo = cases.call(oracle, coords, w, shape, ranges=ranges, bins=bins)
        a, b = cases.padded(hg), cases.padded(ho)
        assert a.shape == b.shape, (trial, nd, shape, method)
        scale = np.max(np.abs(b)) if b.size else 0.0
        assert np.max(np.abs(a - b), initial=0.0) <= 1e-12 * scale + 1e-300, (trial, nd, shape, weighted, method, n, bins, ranges)
        for x, y in zip(eg, eo):
            assert np.array_equal(x, y)
        checked += 1
    ptg.method = None
    assert checked == 400


def test_half_gigabyte_grid_many_slab_passes():
    """402^3 doubles = 520 MB, four times the L2: eight slab passes for rows / global, cross-checked with
    the un-slabbed runs kernel; also the largest 2-D grid the 2^31-cell limit admits in one axis."""
    import torch
    from postpic_b200 import device
    n = 5_000_000
    g = torch.Generator(device='cuda').manual_seed(123)
    c = [torch.rand(n, dtype=torch.float64, device='cuda', generator=g) for _ in range(3)]
    w = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    kw = dict(range=[(0, 1)] * 3, bins=(400, 400, 400), shape=1)
    ref, _ = device.deposit(c, w, method='runs', **kw)
    for m in ('rows', None):
        got, _ = device.deposit(c, w, method=m, **kw)
        assert float((got - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    del ref, got
    torch.cuda.empty_cache()
    # a long, thin 2-D grid: 40000 x 5000 bins = 2e8 cells (1.6 GB)
    kw = dict(range=[(0, 1), (0, 1)], bins=(40000, 5000), shape=2)
    a, _ = device.deposit(c[:2], w, method='runs', **kw)
    b, _ = device.deposit(c[:2], w, method=None, **kw)
    assert float((a - b).abs().max()) <= 1e-12 * float(a.abs().max())
    assert abs(float(b.sum()) - float(w.sum())) <= 1e-12 * float(w.sum())


@pytest.mark.parametrize("method", METHODS)
def test_masked_deposit_equals_compressed_copy(method):
    """ptg_deposit_ex_f64 with a selection mask == depositing the compressed arrays (what postpic's
    filter/compress materialises), for every method, random and cell-ordered input, bool and strided masks."""
    import torch
    from postpic_b200 import device
    n = 3_000_000
    g = torch.Generator(device='cuda').manual_seed(31)
    x = torch.rand(n, dtype=torch.float64, device='cuda', generator=g) * 1.04 - 0.02
    y = torch.rand(n, dtype=torch.float64, device='cuda', generator=g) * 1.04 - 0.02
    w = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    m = torch.rand(n, device='cuda', generator=g) < 0.37
    order = torch.argsort((x.clamp(0, 1) * 150).floor() * 150 + (y.clamp(0, 1) * 150).floor())
    for idx in (None, order):
        sel = (lambda t: t) if idx is None else (lambda t: t[idx].contiguous())
        xs, ys, ws, ms = sel(x), sel(y), sel(w), sel(m)
        for shape in ((0, 1, 2) if method != 'runs_wide' else (0, 1)):
            kw = dict(range=[(0, 1), (0, 1)], bins=(150, 150), shape=shape)
            ref, _ = device.deposit([xs[ms], ys[ms]], ws[ms], method='global', **kw)
            got, _ = device.deposit([xs, ys], ws, method=method, mask=ms, **kw)
            assert float((got - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
            got, _ = device.deposit([xs, ys], None, method=method, mask=ms.to(torch.uint8), **kw)
            ref, _ = device.deposit([xs[ms], ys[ms]], None, method='global', **kw)
            assert float((got - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    # all-false and all-true masks
    kw = dict(range=[(0, 1), (0, 1)], bins=(64, 64), shape=2)
    got, _ = device.deposit([x, y], w, method=method, mask=torch.zeros(n, dtype=torch.bool, device='cuda'), **kw)
    assert not got.any()
    got, _ = device.deposit([x, y], w, method=method, mask=torch.ones(n, dtype=torch.bool, device='cuda'), **kw)
    ref, _ = device.deposit([x, y], w, method='global', **kw)
    assert float((got - ref).abs().max()) <= 1e-12 * float(ref.abs().max())


_BULK_SCRIPT = r'''
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1])
from oracle import ptg_oracle as oracle
from postpic_b200 import device
rng = np.random.default_rng(77)
bad = []
for nd in (1, 2, 3):
    bins = {1: (300,), 2: (40, 30), 3: (12, 10, 8)}[nd]
    rngs = [(0.0, 1.0), (-1.0, 1.0), (0.0, 2.0)][:nd]
    for shape in (0, 1, 2, 3):
        for weighted in (False, True):
            for order in ('random', 'sorted'):
                n = 20000 + 13                                   # whole tiles of 256 plus a ragged tail
                cols = [rng.uniform(lo - 0.05 * (hi - lo), hi + 0.05 * (hi - lo), n) for lo, hi in rngs]
                if order == 'sorted':                            # long runs: the kernel's fast case
                    key = np.zeros(n, dtype=np.int64)
                    for c, (lo, hi), b in zip(cols, rngs, bins):
                        key = key * (b + 8) + np.floor((c - lo) / (hi - lo) * b).astype(np.int64) + 4
                    perm = np.argsort(key, kind='stable')
                    cols = [c[perm] for c in cols]
                w = rng.uniform(0.1, 1.0, n) if weighted else None
                fn = {1: oracle.histogram, 2: oracle.histogram2d, 3: oracle.histogram3d}[nd]
                if nd == 1:
                    want = fn(cols[0], range=rngs[0], bins=bins[0], weights=w, shape=shape)[0].base
                else:
                    want = fn(*cols, weights=w, range=rngs, bins=bins, shape=shape)[0].base
                t = [torch.from_numpy(c).cuda() for c in cols]
                tw = torch.from_numpy(w).cuda() if weighted else None
                grid, plan = device.deposit(t, tw, range=rngs if nd > 1 else rngs[0], bins=bins if nd > 1 else bins[0],
                                            shape=shape, method='runs')
                got = grid.cpu().numpy()
                want = want if want is not None else np.zeros_like(got)
                tol = 1e-12 * max(np.max(np.abs(want)), 1e-300)
                if got.shape != want.shape or np.max(np.abs(got - want)) > tol:
                    bad.append((nd, shape, weighted, order, float(np.max(np.abs(got - want)))))
# several tiles per warp of unordered particles on a grid larger than L2: the case in which a refill of a
# shared-memory slot once overtook the loads that were still reading it (particles with x, y and weight
# from different tiles).  Index-coded cells: every particle must be found exactly once in its own cell.
sys.path.insert(0, sys.argv[1] + '/profiles/diag')
import bulk_runs_diag
nbad = bulk_runs_diag.run(1 << 21, 1)
if nbad:
    bad.append(('index-coded multi-tile', nbad))
print('BAD', bad)
sys.exit(1 if bad else 0)
'''


@pytest.mark.gpu
def test_bulk_copy_runs_kernel_every_instance_against_the_oracle():
    """k_deposit_runs_bulk (cp.async.bulk + mbarrier ring) is the default only where it was measured
    faster; PTG_RUNS_IMPL=b forces it wherever it exists.  Every (ndim, shape, weighted) instance, random
    and cell-ordered input, whole tiles plus a ragged tail, against the oracle -- in a child process because
    the knob is read once per process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PTG_RUNS_IMPL='b')
    r = subprocess.run([sys.executable, '-c', _BULK_SCRIPT, root], env=env, capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_partially_ordered_dump_and_thread_local_method(ptg, oracle):
    """A dump that is cell-ordered in its first 70 % and unordered in the rest: AUTO must not be talked into RUNS by the
    average (the probe asks for ordered blocks, not an ordered mean), and whatever it picks must match the oracle.
    `use_method` overrides the process-wide knob for one thread only."""
    import threading
    import torch
    from postpic_b200 import device, _lib
    rng = np.random.default_rng(17)
    n = 6_000_000
    x, y, w = rng.random(n), rng.random(n), rng.random(n)
    k = int(0.7 * n)
    cell = np.floor(x[:k] * 256) * 256 + np.floor(y[:k] * 256)
    o = np.argsort(cell, kind='stable')
    x[:k], y[:k], w[:k] = x[:k][o], y[:k][o], w[:k][o]
    kw = dict(range=((0, 1), (0, 1)), bins=(256, 256), shape=2)
    got, _ = device.deposit([torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()], torch.from_numpy(w).cuda(), **kw)
    assert _lib.last_probe()['mode'] != 'runs'
    h = oracle.histogram2d(x, y, weights=w, **kw)[0]
    assert np.max(np.abs(got.cpu().numpy() - h.base)) <= 1e-12 * np.max(np.abs(h.base))
    # thread-local method
    seen = {}

    def worker():
        seen['inner'] = ptg._current_method()
    with ptg.use_method('fixed'):
        assert ptg._current_method() == 'fixed'
        t = threading.Thread(target=worker)
        t.start()
        t.join()
        h1 = ptg.histogram2d(x[:200000], y[:200000], weights=w[:200000], **kw)[0]
        h2 = ptg.histogram2d(x[:200000], y[:200000], weights=w[:200000], **kw)[0]
        assert np.array_equal(h1, h2)
    assert seen['inner'] is None and ptg._current_method() is None
    with pytest.raises(ValueError):
        with ptg.use_method('nonsense'):
            pass
