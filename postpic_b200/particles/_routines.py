"""`histogramdd` -- the numpy.histogramdd-like front door of the deposition path.

Re-statement of /root/reference/postpic/particles/_routines.py:39-150 on top of the
CUDA drop-in `_particlestogrid`: same keyword handling (`shape` defaults to 2 and
`shape=None` means the default, :76-79; unknown keywords raise TypeError, :83-84),
same accepted data layouts (:86-102), same default bin counts (:103-107), float64
conversion (:111-113), range completion from the data's min/max (:115-127) and the
1-/2-/3-D dispatch (:129-148).

Beyond the reference: every data/weights array may be a 1-D float64 torch CUDA tensor
(or a (N, D) CUDA tensor); missing range ends are then reduced on the GPU
(ptg_minmax_f64) instead of by np.min/np.max, and nothing crosses PCIe except the
final grid.
"""
import numpy as np

try:
    from collections.abc import Iterable
except ImportError:                                   # pragma: no cover
    from collections import Iterable

from . import _particlestogrid as ptg
from .. import device as _device

particleshapes = ptg.shapes

__all__ = ['histogramdd', 'particleshapes']

_DEFAULT_BINS = {1: [800], 2: [500, 500], 3: [200, 200, 200]}


def _is_tensor(a):
    return type(a).__module__.split('.')[0] == 'torch' and hasattr(a, 'dim')


def _split_columns(data):
    """(N, D) array with N > 3 -> list of D strided column views (_routines.py:86-92)."""
    shp = getattr(data, 'shape', None)
    if shp is not None and len(shp) == 2 and shp[0] > 3:
        return [data[:, i] for i in range(shp[1])]
    return data


def _needs_range(kwrange, ndim):
    """True when at least one range end has to come from the data."""
    for ax in range(ndim):
        for i in (0, 1):
            try:
                if kwrange[ax][i] is None:
                    return True
            except TypeError:
                return True
    return False


def _minmax(arr, cache):
    """(min, max) of one axis, computed once: on the GPU for CUDA tensors (ptg_minmax_f64) and for large
    host arrays (ptg_minmax_host_f64: one pipelined pass over PCIe instead of numpy's two single-threaded
    passes, _routines.py:117-126), with numpy for small host arrays."""
    key = id(arr)
    if key not in cache:
        if _device.is_cuda_tensor(arr):
            cache[key] = _device.minmax(arr)
        elif arr.size >= (1 << 20) and arr.strides[0] % 8 == 0:
            import ctypes
            from .. import _lib
            out = np.empty(2)
            _lib.check(_lib.load().ptg_minmax_host_f64(arr.size, arr.ctypes.data, arr.strides[0] // 8,
                                                       out.ctypes.data,
                                                       -1 if ptg.device_index is None else int(ptg.device_index)))
            cache[key] = (float(out[0]), float(out[1]))
        else:
            if arr.size == 0:
                raise ValueError('zero-size array to reduction operation minimum which has no identity')
            cache[key] = (np.min(arr), np.max(arr))
    return cache[key]


def histogramdd(data, **kwargs):
    '''
    Creates a histogram of the data. Same signature and return values as
    `numpy.histogramdd`, plus the `shape` keyword selecting the particle shape
    (0 NGP, 1 tophat, 2 triangle -- the default --, 3 spline3).  With `shape=0` the
    result equals `numpy.histogramdd`'s.

    Parameters
    ----------
    data: a 1-D array (1-D histogram), a sequence `(datax, datay, dataz)` of 1-D
        arrays (preferred), or a (N, D) array.  numpy arrays or torch CUDA tensors.
    bins: sequence or int -- number of bins per dimension
    range: sequence of (lower, upper) per dimension; missing ends default to the
        data's minimum / maximum
    weights: 1-D array of per-particle weights
    shape: particle shape order or name

    Returns
    -------
    H : ndarray, the histogram;  edges : tuple of D edge arrays
    '''
    kwshape = kwargs.pop('shape', None)
    kwshape = 2 if kwshape is None else kwshape
    kwrange = kwargs.pop('range', None)
    kwweights = kwargs.pop('weights', None)
    kwbins = kwargs.pop('bins', None)
    if len(kwargs) > 0:
        raise TypeError("got an unexpected keyword argument {}'".format(kwargs))

    data = _split_columns(data)
    # a bare 1-D array means a 1-D histogram
    if _is_tensor(data):
        if data.dim() == 1:
            data = (data,)
    elif np.isscalar(data[0]):
        data = (data,)
    if len(data) > 3:
        raise ValueError('Data with len {:} not supported. Maximum is 3D data.'.format(len(data)))
    ndim = len(data)
    if isinstance(kwrange, Iterable) and np.isscalar(kwrange[0]):
        kwrange = (kwrange,)
    if np.isscalar(kwbins):
        kwbins = (kwbins,) * ndim
    if kwbins is None:
        kwbins = _DEFAULT_BINS[ndim]

    on_device = any(_device.is_cuda_tensor(d) for d in data)
    if on_device:
        torch = _device._torch()
        dev = [d for d in data if _device.is_cuda_tensor(d)][0].device
        data = [d if _device.is_cuda_tensor(d) else torch.as_tensor(np.asarray(d, dtype='float64'), device=dev)
                for d in data]
        data = [d if d.dtype == torch.float64 else d.to(torch.float64) for d in data]
        if kwweights is not None:
            if not _device.is_cuda_tensor(kwweights):
                kwweights = torch.as_tensor(np.asarray(kwweights, dtype='float64'), device=dev)
            elif kwweights.dtype != torch.float64:
                kwweights = kwweights.to(torch.float64)
    else:
        data = [np.asarray(d, dtype='float64') for d in data]
        if kwweights is not None:
            kwweights = np.asarray(kwweights, dtype='float64')

    # Large host arrays whose range must first be found are needed twice (min/max, then the deposit):
    # move them to the GPU once -- through the library's threaded pinned staging ring -- when they fit
    # comfortably, so that PCIe is crossed one time instead of two.
    if not on_device and _needs_range(kwrange, ndim) and data[0].size >= (1 << 20):
        try:
            torch = _device._torch()
            free = torch.cuda.mem_get_info()[0] if torch.cuda.is_available() else 0
            need = 8 * data[0].size * (ndim + (kwweights is not None))
            if need < free // 2:
                idx = ptg.device_index
                data = [_device.upload(d, idx) for d in data]
                if kwweights is not None:
                    kwweights = _device.upload(kwweights, idx)
                on_device = True
        except ImportError:                                 # no torch: keep the two-pass host route
            pass

    # complete the ranges: a missing (or None) end is the data's min / max
    ranges = [[None, None] for _ in data]
    mm_cache = {}
    for ax, d in enumerate(data):
        for i in (0, 1):
            try:
                ranges[ax][i] = kwrange[ax][i]
                if ranges[ax][i] is None:
                    raise TypeError
                if not np.isscalar(ranges[ax][i]):
                    raise ValueError('range="{}" not properly formatted.'.format(kwrange))
            except TypeError:
                ranges[ax][i] = _minmax(d, mm_cache)[i]

    if ndim == 1:
        h, xedges = ptg.histogram(data[0], range=ranges[0], bins=kwbins[0],
                                  weights=kwweights, shape=kwshape)
        return h, (xedges,)
    elif ndim == 2:
        h, xedges, yedges = ptg.histogram2d(data[0], data[1], range=ranges, bins=kwbins,
                                            weights=kwweights, shape=kwshape)
        return h, (xedges, yedges)
    else:
        h, xe, ye, ze = ptg.histogram3d(data[0], data[1], data[2], range=ranges, bins=kwbins,
                                        weights=kwweights, shape=kwshape)
        return h, (xe, ye, ze)
