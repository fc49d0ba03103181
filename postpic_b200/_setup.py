"""Host-side set-up shared by every entry point: shape-name lookup, per-axis range
widening, edges and the 1/dx factor -- the few scalar lines that precede the hot loop
in the reference (/root/reference/postpic/particles/_particlestogrid.pyx:29-34 shape
table, :51-60 1-D, :152-169 2-D, :303-328 3-D).  Pure numpy scalar work, no arrays of
particles are touched here.
"""
import numpy as np

# pyx:29-34 -- re-exported as postpic.particles.particleshapes (_routines.py:34)
shapes = [
    [0, 'NGP'],
    [1, 'tophat'],
    [2, 'triangle'],
    [3, 'spline3'],
]


def shape_order(shape):
    """Particle-shape order for an int, bool, float-equal or name spelling.

    The reference dispatches with `shape in shapes[k]` (pyx:66,:77,:92,:110), i.e. an
    equality test against both the number and the name; an unknown shape leaves its
    result array unassigned (a crash).  Here it is a ValueError."""
    for order, name in shapes:
        try:
            if shape in (order, name):
                return order
        except Exception:        # e.g. an ndarray compared with a str
            break
    raise ValueError('unknown particle shape {!r}; choose from {}'.format(shape, shapes))


def axis_setup(rng, bins):
    """(lo, inv, edges) of one axis.

    pyx:52-60: unpack the two range ends (TypeError if range is None), widen a
    degenerate range by one spacing on both sides, edges = linspace, and
    inv = 1.0 / (hi - lo) * bins evaluated left to right."""
    lo, hi = rng
    lo = float(lo)
    hi = float(hi)
    s = np.spacing(hi)
    if np.abs(hi - lo) < s:
        hi += s
        lo -= s
    edges = np.linspace(lo, hi, bins + 1)
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        inv = float(np.float64(1.0) / np.float64(hi - lo) * bins)
    return lo, inv, edges


def as_bins(bins, ndim):
    """C-int conversion of the bins argument (pyx:38 `int bins`, :164-165, :320-322)."""
    if ndim == 1:
        out = [int(bins)]
    else:
        out = [int(bins[a]) for a in range(ndim)]
    for b in out:
        if b < 1:
            raise ValueError('bins must be >= 1, got {}'.format(b))
        if b >= 2 ** 31:
            raise OverflowError('value too large to convert to int')
    return out
