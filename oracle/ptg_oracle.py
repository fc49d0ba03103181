"""Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY -- see ptg_oracle.c.

Same call signatures as the reference module
(/root/reference/postpic/particles/_particlestogrid.pyx:38, :135, :284) so that
tests can swap `oracle._ref._particlestogrid`, this module and the CUDA drop-in
freely.  The arithmetic is in ptg_oracle.c; this file restates only the host-side
set-up: degenerate-range widening (pyx:53-57, :155-163, :307-319), the edges
(np.linspace, pyx:58, :166-167, :323-325) and 1/dx = 1.0/(hi-lo)*bins evaluated
left to right (pyx:60, :168-169, :326-328).
"""
import ctypes
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import build_oracle  # noqa: E402

shapes = [[0, 'NGP'], [1, 'tophat'], [2, 'triangle'], [3, 'spline3']]

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle.build())
        P = ctypes.c_void_p
        _lib.ptg_oracle_deposit.restype = ctypes.c_int64
        _lib.ptg_oracle_deposit.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                            P, P, P, ctypes.c_int64, P, P, P, P]
        _lib.ptg_oracle_accepted_weight.restype = ctypes.c_double
        _lib.ptg_oracle_accepted_weight.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                                    P, P, P, ctypes.c_int64, P, P, P, P]
    return _lib


def shape_id(shape):
    for sid, name in shapes:
        if shape == sid or shape == name:   # same test as `shape in shapes[k]`
            return sid
    raise ValueError('unknown particle shape {!r}'.format(shape))


def axis_setup(rng, bins):
    lo, hi = rng
    lo = float(lo)
    hi = float(hi)
    s = np.spacing(hi)
    if np.abs(hi - lo) < s:
        hi += s
        lo -= s
    edges = np.linspace(lo, hi, bins + 1)
    inv = 1.0 / (hi - lo) * bins
    return lo, inv, edges


def _prep(arrs, weights, ranges, bins, shape):
    sid = shape_id(shape)
    nd = len(arrs)
    for a in arrs:
        if a.dtype != np.float64 or a.ndim != 1:
            raise ValueError('data must be 1-D float64')
    n = len(arrs[0])
    for a, nm in zip(arrs[1:], 'yz'):
        if len(a) != n:
            raise ValueError('datax and data{} must be of equal length'.format(nm))
    if weights is not None and len(weights) < n:
        raise ValueError('weights shorter than data')
    bins = [int(b) for b in bins]
    lo = np.zeros(3)
    inv = np.ones(3)
    edges = []
    for a in range(nd):
        lo[a], inv[a], e = axis_setup(ranges[a], bins[a])
        edges.append(e)
    return sid, n, bins, lo, inv, edges


def deposit(arrs, weights, ranges, bins, shape):
    """Returns (padded grid, edges list, accepted count)."""
    sid, n, bins, lo, inv, edges = _prep(arrs, weights, ranges, bins, shape)
    nd = len(arrs)
    grid = np.zeros([b + 2 * sid for b in bins], dtype=np.float64)
    ptrs = (ctypes.c_void_p * 3)(*[a.ctypes.data for a in arrs] + [None] * (3 - nd))
    strides = (ctypes.c_int64 * 3)(*[a.strides[0] // 8 if n else 1 for a in arrs] + [1] * (3 - nd))
    cb = (ctypes.c_int * 3)(*(bins + [1] * (3 - nd)))
    wptr, ws = (None, 1) if weights is None else (weights.ctypes.data, weights.strides[0] // 8 if n else 1)
    acc = lib().ptg_oracle_deposit(nd, sid, n, ptrs, strides, wptr, ws,
                                   lo.ctypes.data, inv.ctypes.data, cb, grid.ctypes.data)
    if acc < 0:
        raise ValueError('bad arguments')
    return grid, edges, acc


def accepted_weight(arrs, weights, ranges, bins, shape):
    """(sum w, sum |w|) over the particles the acceptance tests let through."""
    sid, n, bins, lo, inv, _ = _prep(arrs, weights, ranges, bins, shape)
    nd = len(arrs)
    ptrs = (ctypes.c_void_p * 3)(*[a.ctypes.data for a in arrs] + [None] * (3 - nd))
    strides = (ctypes.c_int64 * 3)(*[a.strides[0] // 8 if n else 1 for a in arrs] + [1] * (3 - nd))
    cb = (ctypes.c_int * 3)(*(bins + [1] * (3 - nd)))
    wptr, ws = (None, 1) if weights is None else (weights.ctypes.data, weights.strides[0] // 8 if n else 1)
    asum = ctypes.c_double(0.0)
    tot = lib().ptg_oracle_accepted_weight(nd, sid, n, ptrs, strides, wptr, ws,
                                           lo.ctypes.data, inv.ctypes.data, cb, ctypes.byref(asum))
    return tot, asum.value


def _interior(grid, sid, bins):
    return grid[tuple(slice(sid, sid + b) for b in bins)]


def histogram(data, range=None, bins=20, weights=None, shape=0):
    grid, edges, _ = deposit([data], weights, [range], [bins], shape)
    return _interior(grid, shape_id(shape), [int(bins)]), edges[0]


def histogram2d(datax, datay, weights=None, range=None, bins=(20, 20), shape=0):
    grid, edges, _ = deposit([datax, datay], weights, range, bins, shape)
    return (_interior(grid, shape_id(shape), [int(b) for b in bins]),) + tuple(edges)


def histogram3d(datax, datay, dataz, weights=None, range=None, bins=(20, 20, 20), shape=0):
    grid, edges, _ = deposit([datax, datay, dataz], weights, range, bins, shape)
    return (_interior(grid, shape_id(shape), [int(b) for b in bins]),) + tuple(edges)
