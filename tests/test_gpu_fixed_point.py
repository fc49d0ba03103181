"""The 64-bit fixed-point shared-memory tile of method TILED (ptg_deposit_tiled.cu): accuracy contract,
range guard and fall-backs.

Contract (include/ptg_b200.h, PTG_EXACT_ACCUM): a contribution is rounded to a multiple of q = 2^(e-45), 2^e the
power of two above the largest sampled |weight|; per bin |error| <= n_bin * q / 2; integer-valued weights are
exact; weights above 2^e, NaN and inf take the exact path; the tile is merged into the fp64 grid before any
cell could leave the 64-bit range.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KW2 = dict(range=((0.0, 1.0), (0.0, 1.0)), bins=(96, 64))


def _dep(coords, w, method, **kw):
    from postpic_b200 import device
    g, plan = device.deposit(coords, w, method=method, **kw)
    return g


def _cuda(*arrs):
    import torch
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs]


@pytest.mark.parametrize("shape", [0, 1, 2, 3])
def test_integer_weights_are_exact(shape):
    rng = np.random.default_rng(10 + shape)
    n = 3_000_000
    x, y = rng.random(n), rng.random(n)
    w = rng.integers(-7, 9, n).astype(np.float64) * (1 if shape == 0 else 1)
    dx, dy, dw = _cuda(x, y, w)
    if shape == 0:
        ref = _dep([dx, dy], dw, 'global', shape=0, **KW2)
        got = _dep([dx, dy], dw, 'tiled', shape=0, **KW2)
        import torch
        assert torch.equal(got, ref)          # sums of integers: exact in fp64 and in fixed point alike
    else:
        # stencil weights are not integers, but the total deposited weight of integer weights must be conserved to
        # the quantum: |sum error| <= contributions * q / 2 with q = 2^(4 - 45)
        got = _dep([dx, dy], dw, 'tiled', shape=shape, **KW2)
        nw = (shape + 1) ** 2
        assert abs(float(got.sum()) - float(w.sum())) <= n * nw * 2.0 ** -42 + 1e-9


@pytest.mark.parametrize("nd,shape", [(1, 1), (2, 2), (2, 3), (3, 1)])
def test_error_bound_per_bin(oracle, nd, shape):
    """|tiled - oracle| per bin stays below n_bin * q / 2 (+ the fp64 summation noise of the oracle itself)."""
    import torch
    from postpic_b200 import device
    rng = np.random.default_rng(20 + nd + shape)
    n = 2_000_000
    bins = [(1000,), (96, 64), (24, 20, 16)][nd - 1]
    coords = [rng.random(n) for _ in range(nd)]
    w = rng.random(n)
    ranges = [(0.0, 1.0)] * nd
    if nd == 1:
        h = oracle.histogram(coords[0], range=ranges[0], bins=bins[0], weights=w, shape=shape)[0]
    elif nd == 2:
        h = oracle.histogram2d(coords[0], coords[1], weights=w, range=ranges, bins=bins, shape=shape)[0]
    else:
        h = oracle.histogram3d(coords[0], coords[1], coords[2], weights=w, range=ranges, bins=bins, shape=shape)[0]
    ref = h.base if h.base is not None else h
    dc = _cuda(*coords)
    dw = _cuda(w)[0]
    got = _dep(dc, dw, 'tiled', range=ranges if nd > 1 else ranges[0], bins=bins if nd > 1 else bins[0], shape=shape).cpu().numpy()
    cnt = _dep(dc, None, 'global', range=ranges if nd > 1 else ranges[0], bins=bins if nd > 1 else bins[0], shape=0)
    q = 2.0 ** -45                                        # weights in [0, 1): e = 0
    # contributions per bin <= particles in the (S+1)^nd neighbourhood; bound it by the busiest cell
    n_bin = float(cnt.max()) * (shape + 1) ** nd * 2
    err = np.max(np.abs(got - ref))
    assert err <= n_bin * q / 2 + 64 * np.finfo(float).eps * np.max(np.abs(ref))
    assert err <= 1e-12 * np.max(np.abs(ref))
    assert abs(got.sum() - ref.sum()) <= 1e-12 * w.sum()


def test_wide_dynamic_range_and_signs():
    """Weights over 11 decades and both signs: error relative to the LARGEST bin stays within the 1e-12 bar, and the
    absolute error is bounded by the quantum of the largest weight."""
    rng = np.random.default_rng(31)
    n = 4_000_000
    x, y = rng.random(n), rng.random(n)
    w = 10.0 ** rng.uniform(-8, 3, n) * rng.choice([-1.0, 1.0], n)
    dx, dy, dw = _cuda(x, y, w)
    ref = _dep([dx, dy], dw, 'global', shape=2, **KW2)
    got = _dep([dx, dy], dw, 'tiled', shape=2, **KW2)
    err = float((got - ref).abs().max())
    assert err <= 1e-12 * float(ref.abs().max())
    q = 2.0 ** (10 - 45)                                  # max |w| just below 1e3 < 2^10
    assert err <= (n / (96 * 64)) * 9 * 4 * q / 2 + 1e-13 * float(ref.abs().max())


def test_range_guard_many_particles_in_one_cell():
    """5e7 unit-weight particles in ONE cell: sum |w| = 5e7 >> 2^16, so the tile must be merged again and again before
    its 64-bit cells could overflow; the result is exact (integers)."""
    import torch
    n = 50_000_000
    g = torch.Generator(device='cuda').manual_seed(3)
    x = 0.5 + 1e-4 * torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    y = 0.5 + 1e-4 * torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    w = torch.ones(n, dtype=torch.float64, device='cuda')
    for kw in (dict(range=((0.0, 1.0), (0.0, 1.0)), bins=(96, 64)), dict(range=((0.0, 1.0), (0.0, 1.0)), bins=(1024, 1024))):
        got = _dep([x, y], w, 'tiled', shape=0, **kw)
        assert float(got.sum()) == float(n) and float(got.max()) == float(n)
        got = _dep([x, y], w, 'tiled', shape=2, **kw)                   # 9 cells share the 5e7
        assert abs(float(got.sum()) - n) <= 1e-12 * n
        ref = _dep([x, y], w, 'global', shape=2, **kw)
        assert float((got - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    # 1-D, whole grid in the tile, weights of 0.75: same guard in the small-grid kernel
    got = _dep([x], 0.75 * w, 'tiled', range=(0.0, 1.0), bins=1000, shape=1)
    assert abs(float(got.sum()) - 0.75 * n) <= 1e-12 * n


def test_weights_above_the_sampled_cap_take_the_exact_path():
    """The probe samples 16 blocks of 512 particles; a handful of particles elsewhere carry weights 1e6 times larger
    (and one NaN, one inf): they must not be clipped or wrapped."""
    import torch
    rng = np.random.default_rng(41)
    n = 3_000_000
    x, y = rng.random(n), rng.random(n)
    w = rng.random(n)
    big = np.array([1000, 777_777, 1_500_001, 2_222_222])
    w[big] = [1e6, -3e7, 2.5e6, 1e9]
    dx, dy, dw = _cuda(x, y, w)
    ref = _dep([dx, dy], dw, 'global', shape=2, **KW2)
    got = _dep([dx, dy], dw, 'tiled', shape=2, **KW2)
    assert float((got - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    w[12345] = np.inf
    w[2_000_000] = np.nan
    dw = _cuda(w)[0]
    ref = _dep([dx, dy], dw, 'global', shape=2, **KW2)
    got = _dep([dx, dy], dw, 'tiled', shape=2, **KW2)
    assert torch.equal(torch.isfinite(got), torch.isfinite(ref))
    fin = torch.isfinite(ref)
    assert float((got[fin] - ref[fin]).abs().max()) <= 1e-12 * float(ref[fin].abs().max())


def test_zero_and_tiny_weights():
    rng = np.random.default_rng(51)
    n = 1_000_000
    x, y = rng.random(n), rng.random(n)
    dx, dy = _cuda(x, y)
    import torch
    z = torch.zeros(n, dtype=torch.float64, device='cuda')
    got = _dep([dx, dy], z, 'tiled', shape=2, **KW2)
    assert float(got.abs().max()) == 0.0
    tiny = _cuda(rng.random(n) * 1e-290)[0]               # below the supported exponent range: exact path
    ref = _dep([dx, dy], tiny, 'global', shape=2, **KW2)
    got = _dep([dx, dy], tiny, 'tiled', shape=2, **KW2)
    assert float((got - ref).abs().max()) <= 1e-12 * float(ref.abs().max())


def test_exact_flag_keeps_fp64_tile():
    """method 'tiled_exact' (PTG_EXACT_ACCUM): fp64 CAS tile -- single-stream order effects only (~1e-16)."""
    rng = np.random.default_rng(61)
    n = 2_000_000
    x, y = rng.random(n), rng.random(n)
    w = rng.random(n)
    dx, dy, dw = _cuda(x, y, w)
    ref = _dep([dx, dy], dw, 'global', shape=3, **KW2)
    got = _dep([dx, dy], dw, 'tiled_exact', shape=3, **KW2)
    assert float((got - ref).abs().max()) <= 1e-14 * float(ref.abs().max())


@pytest.mark.parametrize("nd,shape,weighted", [(1, 0, False), (1, 1, True), (1, 3, True), (2, 2, True), (2, 0, True), (3, 1, False)])
def test_single_launch_small_path(oracle, nd, shape, weighted):
    """AUTO on small calls (n < 2^21, <= 4096 padded cells): one kernel, no memset, no probe; parity with the oracle,
    bit-reproducible, correct both when the grid is cleared by the call and when it is accumulated into."""
    import torch
    from postpic_b200 import device
    rng = np.random.default_rng(70 + nd + shape)
    n = 700_001
    bins = [(1000,), (50, 40), (12, 10, 9)][nd - 1]
    ranges = [(0.0, 1.0)] * nd
    coords = [rng.uniform(-0.05, 1.05, n) for _ in range(nd)]
    w = rng.random(n) - 0.3 if weighted else None
    if weighted:
        w[1000] = np.nan
    if nd == 1:
        h = oracle.histogram(coords[0], range=ranges[0], bins=bins[0], weights=w, shape=shape)[0]
    elif nd == 2:
        h = oracle.histogram2d(coords[0], coords[1], weights=w, range=ranges, bins=bins, shape=shape)[0]
    else:
        h = oracle.histogram3d(coords[0], coords[1], coords[2], weights=w, range=ranges, bins=bins, shape=shape)[0]
    ref = h.base if h.base is not None else h
    dc = [torch.from_numpy(c).cuda() for c in coords]
    dw = torch.from_numpy(w).cuda() if weighted else None
    kw = dict(range=ranges if nd > 1 else ranges[0], bins=bins if nd > 1 else bins[0], shape=shape)
    g1, plan = device.deposit(dc, dw, **kw)
    g2, _ = device.deposit(dc, dw, **kw)
    assert torch.equal(g1, g2) or bool(torch.isnan(g1).any())
    got = g1.cpu().numpy()
    assert np.array_equal(np.isfinite(got), np.isfinite(ref))
    fin = np.isfinite(ref)
    if shape == 0 and not weighted:
        assert np.array_equal(got, ref)
    else:
        assert np.max(np.abs(got[fin] - ref[fin])) <= 1e-12 * np.max(np.abs(ref[fin]))
    # accumulate mode (zero=False): twice the particles = twice the histogram
    acc = torch.zeros_like(g1)
    device.deposit_into(acc, plan, dc, dw, zero=False)
    device.deposit_into(acc, plan, dc, dw, zero=False)
    a = acc.cpu().numpy()
    assert np.max(np.abs(a[fin] - 2 * ref[fin])) <= 2e-12 * np.max(np.abs(ref[fin]))
