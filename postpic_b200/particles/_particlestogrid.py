"""Drop-in for postpic's only native module, `postpic.particles._particlestogrid`.

Same names, argument order, defaults, return structure and error behaviour as
/root/reference/postpic/particles/_particlestogrid.pyx:
    shapes                                                            :29-34
    histogram(data, range=None, bins=20, weights=None, shape=0)       :38-130
    histogram2d(datax, datay, weights=None, range=None, bins=(20,20), shape=0)       :135-278
    histogram3d(datax, datay, dataz, weights=None, range=None, bins=(20,20,20), shape=0)  :284-465
but the deposition itself runs in hand-written sm_100a CUDA kernels behind the C ABI
of include/ptg_b200.h.  There is no CPU path: without the built library or without a
B200 these functions raise.

Inputs may be numpy arrays (host: streamed to the GPU in chunks, PCIe copy overlapped
with the deposit) or 1-D float64 torch CUDA tensors (device-resident: zero-copy).
Like the reference, the returned histogram is a numpy *view* of the ghost-padded
accumulator (`h.base` holds the ghost cells; the reference's tests read it), it is
float64, owned by the caller and may be modified in place (`h /= dV`,
particles.py:1102).
"""
import ctypes

import numpy as np

from .. import _lib
from .._setup import shapes, shape_order, axis_setup, as_bins     # noqa: F401  (shapes is API)
from .. import device as _device

# deposition method for every call of this module: None/'auto', 'global', 'rows', ...
# (a module attribute rather than an argument so the reference signatures stay untouched)
method = None
# GPU used for numpy inputs; None = the process's current CUDA device
device_index = None

# per-thread override of `method` (the module attribute above is process-wide): `with use_method('fixed'): ...`
import contextlib as _contextlib
import threading as _threading

_local = _threading.local()
_UNSET = object()


@_contextlib.contextmanager
def use_method(name):
    """Deposition method for the calls of THIS thread inside the block (thread-safe, unlike the module attribute)."""
    if name not in _lib.METHODS:
        raise ValueError('unknown deposition method {!r}'.format(name))
    prev = getattr(_local, 'method', _UNSET)
    _local.method = name
    try:
        yield
    finally:
        _local.method = prev


def _current_method():
    m = getattr(_local, 'method', _UNSET)
    return method if m is _UNSET else m


def _host_array(a, name):
    if not isinstance(a, np.ndarray):
        raise TypeError("Argument '{}' has incorrect type (expected numpy.ndarray, got {})"
                        .format(name, type(a).__name__))
    if a.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'double_t' but got '{}'".format(a.dtype.name))
    if a.ndim != 1:
        raise ValueError('Buffer has wrong number of dimensions (expected 1, got {})'.format(a.ndim))
    return a


def _element_stride(a):
    """(array, stride in doubles) for the C ABI, which gathers strided / reversed host views into its
    pinned staging ring itself; only byte strides that are not a multiple of 8 need a copy first."""
    if a.shape[0] <= 1:
        return np.ascontiguousarray(a), 1
    if a.strides[0] % 8 == 0:
        return a, a.strides[0] // 8
    return np.ascontiguousarray(a), 1


def _deposit(coords, names, weights, ranges, bins, shape):
    ndim = len(coords)
    on_device = _device.is_cuda_tensor(coords[0])
    if on_device:
        if not all(_device.is_cuda_tensor(c) for c in coords) or \
                (weights is not None and not _device.is_cuda_tensor(weights)):
            raise TypeError('either all inputs are CUDA tensors or all are numpy arrays')
    else:
        coords = [_host_array(c, nm) for c, nm in zip(coords, names)]
        if weights is not None:
            weights = _host_array(weights, 'weights')
    n = len(coords[0])
    for c, nm in zip(coords[1:], 'yz'):
        if len(c) != n:                                          # pyx:150-151, :299-302
            raise ValueError('datax and data{} must be of equal length'.format(nm))
    order = shape_order(shape)
    cbins = as_bins(bins, ndim)
    plan = _device.Plan([ranges[a] for a in range(ndim)] if ndim > 1 else [ranges], cbins, order)
    if weights is not None and len(weights) < n:
        # the reference reads past the end of a short weights buffer (undefined behaviour)
        raise ValueError('weights must have at least as many elements as data')

    if on_device:
        torch = _device._torch()
        grid_t = torch.empty(plan.dims, dtype=torch.float64, device=coords[0].device)
        _device.deposit_into(grid_t, plan, coords, weights, method=_current_method(), zero=True)
        grid = grid_t.cpu().numpy()
    else:
        lib = _lib.load()
        pairs = [_element_stride(c) for c in coords]
        coords = [p[0] for p in pairs]
        strides = [p[1] for p in pairs] + [1] * (3 - ndim)
        wstride = 1
        if weights is not None:
            weights, wstride = _element_stride(weights)
        cstrides = (ctypes.c_int64 * 4)(*(strides + [wstride]))
        grid = np.empty(plan.dims, dtype=np.float64)
        ptr = [c.ctypes.data if n else None for c in coords] + [None] * (3 - ndim)
        wptr = weights.ctypes.data if (weights is not None and n) else None
        _lib.check(lib.ptg_histogram_host_f64(
            ndim, plan.shape, n, ptr[0], ptr[1], ptr[2], wptr, cstrides,
            plan.lo.ctypes.data, plan.inv.ctypes.data, plan.cbins, grid.ctypes.data,
            _lib.METHODS[_current_method()], -1 if device_index is None else int(device_index)))
    return grid[plan.interior], plan.edges


def histogram(data, range=None, bins=20, weights=None, shape=0):
    """1-D deposit; mimics numpy.histogram plus `shape` (pyx:38-130). Returns (h, edges)."""
    h, edges = _deposit([data], ['data'], weights, range, bins, shape)
    return h, edges[0]


def histogram2d(datax, datay, weights=None, range=None, bins=(20, 20), shape=0):
    """2-D deposit; mimics numpy.histogram2d plus `shape` (pyx:135-278). Returns (h, xedges, yedges)."""
    h, edges = _deposit([datax, datay], ['datax', 'datay'], weights, range, bins, shape)
    return h, edges[0], edges[1]


def histogram3d(datax, datay, dataz, weights=None, range=None, bins=(20, 20, 20), shape=0):
    """3-D deposit (pyx:284-465). Returns (h, xedges, yedges, zedges)."""
    h, edges = _deposit([datax, datay, dataz], ['datax', 'datay', 'dataz'], weights, range, bins, shape)
    return h, edges[0], edges[1], edges[2]
