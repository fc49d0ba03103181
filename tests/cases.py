"""Seeded input generators shared by the golden-vector script and the tests."""
import numpy as np

RANGES = [(0.0, 1.0), (0.02, 1.93), (-0.4, 2.6)]
BINS = [37, 23, 11]


def cloud(seed, n, ndim, weighted, spill=0.1, integer_weights=False):
    """n particles; coordinates overshoot the ranges of RANGES by `spill` on both sides so
    that the boundary rules are exercised."""
    rng = np.random.default_rng(seed)
    coords = []
    for a in range(ndim):
        lo, hi = RANGES[a]
        span = hi - lo
        coords.append(rng.uniform(lo - spill * span, hi + spill * span, n))
    w = None
    if weighted:
        w = rng.integers(1, 8, n).astype(np.float64) if integer_weights else rng.uniform(0.0, 1.0, n)
    return coords, w


def call(mod, coords, w, shape, ranges=None, bins=None):
    """Call histogram / histogram2d / histogram3d of any of the three interchangeable
    modules (reference, oracle, CUDA drop-in); returns (h, [edges...])."""
    nd = len(coords)
    ranges = RANGES[:nd] if ranges is None else ranges
    bins = BINS[:nd] if bins is None else bins
    if nd == 1:
        h, e = mod.histogram(coords[0], range=ranges[0], bins=bins[0], weights=w, shape=shape)
        return h, [e]
    if nd == 2:
        h, ex, ey = mod.histogram2d(coords[0], coords[1], weights=w, range=ranges, bins=bins, shape=shape)
        return h, [ex, ey]
    h, ex, ey, ez = mod.histogram3d(coords[0], coords[1], coords[2], weights=w, range=ranges, bins=bins,
                                    shape=shape)
    return h, [ex, ey, ez]


def padded(h):
    """The ghost-padded accumulator behind a returned histogram view."""
    return h.base if h.base is not None else h


# known-answer vectors (positions from examples/particleshapedemo.py:63-65)
KNOWN_1D_POS = np.array([4.5, 9.75, 15.0, 20.25])
KNOWN_1D = {
    0: {4: 1, 9: 1, 15: 1, 20: 1},
    1: {4: 1.0, 9: 0.75, 10: 0.25, 14: 0.5, 15: 0.5, 19: 0.25, 20: 0.75},
    2: {3: 0.125, 4: 0.75, 5: 0.125, 8: 0.03125, 9: 0.6875, 10: 0.28125, 14: 0.5, 15: 0.5, 19: 0.28125,
        20: 0.6875, 21: 0.03125},
    3: {3: 1 / 6, 4: 2 / 3, 5: 1 / 6, 8: 0.0703125, 9: 0.6119791666666666, 10: 0.3151041666666667,
        11: 0.0026041666666666665, 13: 0.020833333333333332, 14: 0.4791666666666667, 15: 0.4791666666666667,
        16: 0.020833333333333332, 18: 0.0026041666666666665, 19: 0.3151041666666667, 20: 0.6119791666666666,
        21: 0.0703125},
}
KNOWN_2D_BLOCK = np.array([[0.045, 0.41625, 0.10125], [0.11, 1.0175, 0.2475], [0.005, 0.04625, 0.01125]])

# boundary probes: 1-D, range (0,1), bins 10, one particle at `value`;
# expected (sum over padded array, sum over returned bins) for shapes 0,1,2,3
BOUNDARY_PROBES = {
    -0.04: [(0, 0), (1, 0.1), (1, 0.02), (1, 0.2213)],
    -0.06: [(0, 0), (1, -0.1), (1, -0.28), (1, 0.122)],
    -0.14: [(0, 0), (1, -0.9), (0, 0), (1, 0.3647)],
    -0.16: [(0, 0), (0, 0), (0, 0), (0, 0)],
    1.04: [(0, 0), (1, 0.1), (1, 0.18), (1, 0.2213)],
    1.06: [(0, 0), (0, 0), (1, 0.08), (0, 0)],
}
