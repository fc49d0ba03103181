"""Device-resident face of the deposition library.

`deposit()` takes 1-D float64 CUDA tensors (zero-copy, `tensor.data_ptr()` goes
straight into the C ABI) and returns the ghost-padded grid as a CUDA tensor, so a
caller that already has its particles in HBM never crosses PCIe.  PyTorch is used for
device memory, streams and (in `postpic_b200.parallel`) torch.distributed only; all
arithmetic happens in libptg_b200.so.

Replaces the loops of /root/reference/postpic/particles/_particlestogrid.pyx
(:70-129, :181-276, :341-463); set-up arithmetic in `_setup.py`.
"""
import ctypes

import numpy as np

from . import _lib
from ._setup import shape_order, axis_setup, as_bins


def _torch():
    import torch
    return torch


def is_cuda_tensor(a):
    return type(a).__module__.split('.')[0] == 'torch' and getattr(a, 'is_cuda', False)


def _check_tensor(t, name):
    torch = _torch()
    if t.dtype != torch.float64:
        raise ValueError("Buffer dtype mismatch, expected 'double_t' but got '{}' ({})".format(t.dtype, name))
    if t.dim() != 1:
        raise ValueError('Buffer has wrong number of dimensions (expected 1, got {})'.format(t.dim()))
    if t.numel() > 1 and t.stride(0) < 1:
        raise ValueError('{}: negative or zero strides are not supported on device tensors'.format(name))


class Plan(object):
    """Everything about a histogram except the particles: shape order, bins, lo/inv/edges."""

    def __init__(self, ranges, bins, shape):
        self.ndim = len(ranges)
        if not 1 <= self.ndim <= 3:
            raise ValueError('only 1, 2 or 3 axes are supported')
        self.shape = shape_order(shape)
        self.bins = [int(b) for b in bins]
        if len(self.bins) != self.ndim:
            raise ValueError('need one bin count per axis')
        self.lo = np.zeros(3)
        self.inv = np.ones(3)
        self.edges = []
        for a in range(self.ndim):
            self.lo[a], self.inv[a], e = axis_setup(ranges[a], self.bins[a])
            self.edges.append(e)
        self.dims = [b + 2 * self.shape for b in self.bins]
        self.cells = int(np.prod(self.dims, dtype=np.int64))
        self.cbins = (ctypes.c_int * 3)(*(self.bins + [1] * (3 - self.ndim)))
        self.interior = tuple(slice(self.shape, self.shape + b) for b in self.bins)

    def cell_volume(self):
        """prod(edge[1] - edge[0]) -- the dV of particles.py:1101."""
        return float(np.prod([e[1] - e[0] for e in self.edges]))


def deposit_into(grid, plan, coords, weights=None, method=None, zero=True, stream=None, fused=None, mask=None):
    """Deposit CUDA tensors `coords` (+ optional `weights`) into the padded CUDA tensor `grid`.

    Asynchronous on `stream` (default: torch's current stream on grid.device).
    `fused` = dict(kind=, px=, py=, pz=, mass=[, c=, elementary_charge=]) multiplies every particle's
    weight by a kinetic scalar of its momentum inside the deposit kernel (ptg_deposit_fused_f64):
    `weights='Ekin_MeV'` of createField without materialising a weights array.
    `mask` = bool / uint8 CUDA tensor: only particles with a non-zero entry are deposited
    (MultiSpecies.filter as predicated deposition, no compressed copies)."""
    torch = _torch()
    lib = _lib.load()
    if len(coords) != plan.ndim:
        raise ValueError('expected {} coordinate arrays'.format(plan.ndim))
    n = coords[0].numel()
    for t, nm in zip(coords, 'xyz'):
        _check_tensor(t, 'data' + nm)
        if t.numel() != n:
            raise ValueError('datax and data{} must be of equal length'.format(nm))
        if t.device != grid.device:
            raise ValueError('all tensors must live on the same device')
    if weights is not None:
        _check_tensor(weights, 'weights')
        if weights.numel() < n:
            raise ValueError('weights must have at least as many elements as data')
        if weights.device != grid.device:
            raise ValueError('all tensors must live on the same device')
    if grid.dtype != torch.float64 or not grid.is_contiguous() or grid.numel() != plan.cells:
        raise ValueError('grid must be a contiguous float64 tensor of {} elements'.format(plan.cells))
    ptr = [t.data_ptr() if n else None for t in coords] + [None] * (3 - plan.ndim)
    strides = [t.stride(0) if n > 1 else 1 for t in coords] + [1] * (3 - plan.ndim)
    ptr.append(weights.data_ptr() if (weights is not None and n) else None)
    strides.append(weights.stride(0) if (weights is not None and n > 1) else 1)
    cstr = (ctypes.c_int64 * 4)(*strides)
    flags = _lib.METHODS[method] | (_lib.ZERO_GRID if zero else 0)
    if fused is not None:
        q = [fused['px'], fused['py'], fused['pz']]
        for t, nm in zip(q, ('px', 'py', 'pz')):
            _check_tensor(t, nm)
            if t.numel() < n or t.device != grid.device:
                raise ValueError('fused momentum arrays must match the coordinates')
        qstr = (ctypes.c_int64 * 3)(*[t.stride(0) if n > 1 else 1 for t in q])
    if mask is not None:
        if mask.dtype not in (torch.bool, torch.uint8) or mask.dim() != 1 or mask.numel() < n or \
                mask.device != grid.device:
            raise ValueError('mask must be a 1-D bool / uint8 CUDA tensor with one entry per particle')
        args = _lib.DepositArgs()
        args.struct_size = ctypes.sizeof(_lib.DepositArgs)
        args.ndim, args.shape, args.n = plan.ndim, plan.shape, n
        args.x, args.y, args.z, args.w = ptr[0], ptr[1], ptr[2], ptr[3]
        args.strides = ctypes.cast(cstr, ctypes.c_void_p)
        args.lo, args.inv = plan.lo.ctypes.data, plan.inv.ctypes.data
        args.bins = ctypes.cast(plan.cbins, ctypes.c_void_p)
        args.grid, args.flags = grid.data_ptr(), flags
        args.mask, args.mask_stride = (mask.data_ptr() if n else None), (mask.stride(0) if n > 1 else 1)
        if fused is not None:
            args.fuse_kind = WEIGHT_KINDS[fused['kind']]
            args.px, args.py, args.pz = (t.data_ptr() if n else None for t in q)
            args.pstrides = ctypes.cast(qstr, ctypes.c_void_p)
            args.mass = float(fused['mass'])
            args.c = float(fused.get('c', 299792458.0))
            args.elementary_charge = float(fused.get('elementary_charge', 1.602176634e-19))
        with torch.cuda.device(grid.device):
            args.stream = torch.cuda.current_stream().cuda_stream if stream is None else stream
            _lib.check(lib.ptg_deposit_ex_f64(ctypes.byref(args)))
        return grid
    with torch.cuda.device(grid.device):
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        if fused is None:
            _lib.check(lib.ptg_deposit_f64(plan.ndim, plan.shape, n, ptr[0], ptr[1], ptr[2], ptr[3], cstr,
                                           plan.lo.ctypes.data, plan.inv.ctypes.data, plan.cbins,
                                           grid.data_ptr(), flags, None, 0, st))
        else:
            _lib.check(lib.ptg_deposit_fused_f64(
                plan.ndim, plan.shape, n, ptr[0], ptr[1], ptr[2], ptr[3], cstr, plan.lo.ctypes.data,
                plan.inv.ctypes.data, plan.cbins, grid.data_ptr(), flags, WEIGHT_KINDS[fused['kind']],
                q[0].data_ptr() if n else None, q[1].data_ptr() if n else None, q[2].data_ptr() if n else None, qstr,
                float(fused['mass']), float(fused.get('c', 299792458.0)),
                float(fused.get('elementary_charge', 1.602176634e-19)), st))
    return grid


def deposit_host_into(grid, plan, coords, weights=None, method=None, zero=True):
    """Deposit HOST numpy arrays (float64, any element stride, pageable or pinned) into the padded CUDA tensor
    `grid` through the library's chunked H2D pipeline (ptg_deposit_host_f64): the copy of chunk k+1 overlaps
    the deposit of chunk k and the result stays in HBM.  Synchronous."""
    torch = _torch()
    lib = _lib.load()
    if len(coords) != plan.ndim:
        raise ValueError('expected {} coordinate arrays'.format(plan.ndim))
    if grid.dtype != torch.float64 or not grid.is_contiguous() or grid.numel() != plan.cells or not grid.is_cuda:
        raise ValueError('grid must be a contiguous float64 CUDA tensor of {} elements'.format(plan.cells))
    arrs = []
    for a, nm in zip(list(coords) + [weights], ['datax', 'datay', 'dataz'][:plan.ndim] + ['weights']):
        if a is None:
            arrs.append((None, 1))
            continue
        a = np.asarray(a)
        if a.dtype != np.float64 or a.ndim != 1:
            raise ValueError('{}: 1-D float64 array expected'.format(nm))
        if a.shape[0] > 1 and a.strides[0] % 8:
            a = np.ascontiguousarray(a)
        arrs.append((a, a.strides[0] // 8 if a.shape[0] > 1 else 1))
    n = arrs[0][0].shape[0]
    for a, _ in arrs[:plan.ndim]:
        if a.shape[0] != n:
            raise ValueError('datax and datay must be of equal length')
    w, wstride = arrs[-1]
    if w is not None and w.shape[0] < n:
        raise ValueError('weights must have at least as many elements as data')
    ptr = [a.ctypes.data if n else None for a, _ in arrs[:plan.ndim]] + [None] * (3 - plan.ndim)
    strides = [st for _, st in arrs[:plan.ndim]] + [1] * (3 - plan.ndim) + [wstride]
    cstr = (ctypes.c_int64 * 4)(*strides)
    flags = _lib.METHODS[method] | (_lib.ZERO_GRID if zero else 0)
    _lib.check(lib.ptg_deposit_host_f64(plan.ndim, plan.shape, n, ptr[0], ptr[1], ptr[2],
                                        w.ctypes.data if (w is not None and n) else None, cstr,
                                        plan.lo.ctypes.data, plan.inv.ctypes.data, plan.cbins,
                                        grid.data_ptr(), flags, grid.device.index))
    return grid


def deposit(coords, weights=None, range=None, bins=None, shape=0, method=None, fused=None, mask=None):
    """-> (padded grid CUDA tensor, plan).  `plan.interior` slices out the histogram."""
    torch = _torch()
    coords = list(coords)
    nd = len(coords)
    if nd == 1 and np.isscalar(range[0]):
        range = [range]
    plan = Plan(range, as_bins(bins[0] if (nd == 1 and np.ndim(bins)) else bins, nd), shape)
    grid = torch.empty(plan.dims, dtype=torch.float64, device=coords[0].device)
    deposit_into(grid, plan, coords, weights, method=method, zero=True, fused=fused, mask=mask)
    return grid, plan


def minmax(t):
    """(min, max) of a 1-D float64 CUDA tensor as Python floats (NaN propagates like numpy)."""
    torch = _torch()
    _check_tensor(t, 'data')
    if t.numel() == 0:
        raise ValueError('zero-size array to reduction operation minimum which has no identity')
    out = torch.empty(2, dtype=torch.float64, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(_lib.load().ptg_minmax_f64(t.numel(), t.data_ptr(), t.stride(0) if t.numel() > 1 else 1,
                                              out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    lo, hi = out.tolist()
    return lo, hi


def scale_(grid, factor):
    """grid *= factor in place on the device (the h /= dV of particles.py:1102 with factor = 1/dV)."""
    torch = _torch()
    with torch.cuda.device(grid.device):
        _lib.check(_lib.load().ptg_scale_f64(grid.data_ptr(), grid.numel(), float(factor),
                                             torch.cuda.current_stream().cuda_stream))
    return grid


def divide_(grid, divisor):
    """grid /= divisor in place on the device with a true IEEE division: the h /= dV of particles.py:1102, bit for bit."""
    torch = _torch()
    with torch.cuda.device(grid.device):
        _lib.check(_lib.load().ptg_divide_f64(grid.data_ptr(), grid.numel(), float(divisor),
                                              torch.cuda.current_stream().cuda_stream))
    return grid


WEIGHT_KINDS = {'gamma_m1': 1, 'gamma': 2, 'Ekin': 3, 'Ekin_MeV': 4, 'Ekin_keV': 5, 'p': 6}

# postpic's default scalar context (scalarproperties.py:159-222) as the enumerated kinds PTG_S_* of ptg_b200.h
SCALAR_KINDS = dict(WEIGHT_KINDS)
SCALAR_KINDS.update({
    'x': 7, 'y': 8, 'z': 9, 'px': 10, 'py': 11, 'pz': 12, 'weight': 13, 'w': 13, 'mass': 14, 'm': 14,
    'charge': 15, 'q': 15, 'time': 16, 't': 16, 'mass_u': 17, 'm_u': 17, 'charge_e': 18, 'q_e': 18,
    'x_um': 19, 'y_um': 20, 'z_um': 21, '_np2': 22, 'gamma_m': 23, 'beta': 24, 'v': 25, 'vx': 26, 'vy': 27,
    'vz': 28, 'betax': 29, 'betay': 30, 'betaz': 31, 'Eruhe': 32, 'Ekin_MeV_amu': 33, 'Ekin_MeV_qm': 34,
    'Ekin_keV_amu': 35, 'Ekin_keV_qm': 36, 'angle_xaxis': 37, 'angle_yaxis': 38, 'angle_zaxis': 39,
    'angle_xy': 40, 'angle_yx': 41, 'angle_yz': 42, 'angle_zy': 43, 'angle_zx': 44, 'angle_xz': 45,
    'r_xy': 46, 'r_yz': 47, 'r_zx': 48, 'r_xyz': 49})
_NEEDS_POS = {7: 'x', 8: 'y', 9: 'z', 19: 'x', 20: 'y', 21: 'z', 46: 'xy', 47: 'yz', 48: 'zx', 49: 'xyz'}
_NO_ARRAYS = (13, 14, 15, 16, 17, 18, 32)


def scalar_arrays_needed(kind):
    """Names of the atomic arrays scalar `kind` reads ('weight' is optional: a constant works too)."""
    if kind in _NEEDS_POS:
        return list(_NEEDS_POS[kind])
    if kind in _NO_ARRAYS:
        return []
    return ['px', 'py', 'pz']


def _species_args(species, n, device):
    """ptg_deposit_args with the atomic arrays / constants of one species filled in.  `species`: dict with CUDA
    tensors x, y, z, px, py, pz (those that exist), weight (tensor or number), mass, charge, time (numbers)."""
    torch = _torch()
    args = _lib.DepositArgs()
    args.struct_size = ctypes.sizeof(_lib.DepositArgs)
    args.n = n
    keep = []

    def arr(name):
        t = species.get(name)
        if t is None or not is_cuda_tensor(t):
            return None, 1
        if t.dtype != torch.float64:
            t = t.to(torch.float64)
            keep.append(t)
        _check_tensor(t, name)
        if t.numel() < n or t.device != device:
            raise ValueError('{}: wrong length or device'.format(name))
        return (t.data_ptr() if n else None), (t.stride(0) if n > 1 else 1)
    pos = [arr(k) for k in 'xyz']
    mom = [arr(k) for k in ('px', 'py', 'pz')]
    args.pos_x, args.pos_y, args.pos_z = (p[0] for p in pos)
    pstr = (ctypes.c_int64 * 3)(*[p[1] for p in pos])
    args.pos_strides = ctypes.cast(pstr, ctypes.c_void_p)
    args.px, args.py, args.pz = (p[0] for p in mom)
    qstr = (ctypes.c_int64 * 3)(*[p[1] for p in mom])
    args.pstrides = ctypes.cast(qstr, ctypes.c_void_p)
    wt = species.get('weight', 1.0)
    wptr, wstride = arr('weight')
    args.w = wptr
    args.weight_const = 1.0 if wptr is not None else float(wt)
    sstr = (ctypes.c_int64 * 4)(1, 1, 1, wstride)
    args.strides = ctypes.cast(sstr, ctypes.c_void_p)
    args.mass = float(species.get('mass', 0.0) or 0.0)
    args.charge = float(species.get('charge', 0.0) or 0.0)
    args.time = float(species.get('time', 0.0) or 0.0)
    args.c = float(species.get('c', 299792458.0))
    args.elementary_charge = float(species.get('elementary_charge', 1.602176634e-19))
    args.atomic_mass = float(species.get('atomic_mass', 1.66053906660e-27))
    keep += [pstr, qstr, sstr]
    return args, keep


def scalar(name, species, n, device):
    """The per-particle scalar `name` of postpic's default scalar context as a CUDA tensor, evaluated by the
    library (ptg_scalar_f64: the device function the fused deposit kernels call)."""
    torch = _torch()
    args, keep = _species_args(species, n, device)
    out = torch.empty(n, dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        args.stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.load().ptg_scalar_f64(ctypes.byref(args), SCALAR_KINDS[name], out.data_ptr()))
    return out


def deposit_scalars_into(grid, plan, species, n, axes, weight_expr=None, method=None, zero=True, mask=None, stream=None):
    """createField's inner step in ONE pass over a device-resident species: axis a of every particle is the scalar
    `axes[a]`, its weight is weight * `weight_expr` (None: weight alone), both evaluated inside the deposit kernel
    from the atomic arrays in `species` -- no N-sized temporary is written."""
    torch = _torch()
    if len(axes) != plan.ndim:
        raise ValueError('expected {} axis scalars'.format(plan.ndim))
    if grid.dtype != torch.float64 or not grid.is_contiguous() or grid.numel() != plan.cells:
        raise ValueError('grid must be a contiguous float64 tensor of {} elements'.format(plan.cells))
    args, keep = _species_args(species, n, grid.device)
    args.ndim, args.shape = plan.ndim, plan.shape
    for a, nm in enumerate(axes):
        args.axis_kind[a] = SCALAR_KINDS[nm]
    args.fuse_kind = SCALAR_KINDS[weight_expr] if weight_expr is not None else 0
    args.lo, args.inv = plan.lo.ctypes.data, plan.inv.ctypes.data
    args.bins = ctypes.cast(plan.cbins, ctypes.c_void_p)
    args.grid = grid.data_ptr()
    args.flags = _lib.METHODS[method] | (_lib.ZERO_GRID if zero else 0)
    if mask is not None:
        if mask.dtype not in (torch.bool, torch.uint8) or mask.dim() != 1 or mask.numel() < n or mask.device != grid.device:
            raise ValueError('mask must be a 1-D bool / uint8 CUDA tensor with one entry per particle')
        args.mask, args.mask_stride = (mask.data_ptr() if n else None), (mask.stride(0) if n > 1 else 1)
    with torch.cuda.device(grid.device):
        args.stream = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _lib.check(_lib.load().ptg_deposit_ex_f64(ctypes.byref(args)))
    return grid


def kinetic_weights(kind, px, py, pz, weight=None, mass=None, c=299792458.0, elementary_charge=1.602176634e-19):
    """weight * <kind>(px, py, pz) in one pass on the device (kind: gamma_m1, gamma, Ekin, Ekin_MeV,
    Ekin_keV, p) -- the `weight * (expr)` of particles.py:1097 without numexpr's temporaries."""
    torch = _torch()
    for t, nm in ((px, 'px'), (py, 'py'), (pz, 'pz')):
        _check_tensor(t, nm)
    if weight is not None:
        _check_tensor(weight, 'weight')
    n = px.numel()
    out = torch.empty(n, dtype=torch.float64, device=px.device)
    strides = (ctypes.c_int64 * 4)(*[t.stride(0) if n > 1 else 1 for t in (px, py, pz)],
                                   weight.stride(0) if (weight is not None and n > 1) else 1)
    with torch.cuda.device(px.device):
        _lib.check(_lib.load().ptg_weights_f64(
            WEIGHT_KINDS[kind], n, px.data_ptr(), py.data_ptr(), pz.data_ptr(),
            weight.data_ptr() if weight is not None else None, strides, float(mass if mass is not None else 1.0),
            float(c), float(elementary_charge), out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return out


def upload_into(dst, offset, arr):
    """dst[offset:offset+len(arr)] = arr for a host float64 array (any element stride, pageable or pinned) and a
    contiguous float64 CUDA tensor `dst`, through the library's threaded pinned staging ring."""
    a = np.asarray(arr, dtype=np.float64)
    if a.ndim != 1:
        raise ValueError('1-D array expected')
    if a.size > 1 and a.strides[0] % 8:
        a = np.ascontiguousarray(a)
    if offset < 0 or offset + a.size > dst.numel() or not dst.is_contiguous():
        raise ValueError('upload_into: destination range out of bounds')
    if a.size:
        _lib.check(_lib.load().ptg_upload_host_f64(a.size, a.ctypes.data, a.strides[0] // 8 if a.size > 1 else 1,
                                                   dst.data_ptr() + 8 * int(offset), dst.device.index))
    return dst


def upload(arr, device_index=None):
    """Host float64 array (any stride, pageable or pinned) -> contiguous CUDA tensor through the library's
    threaded pinned staging ring (about 4x faster than torch's pageable copy)."""
    torch = _torch()
    a = np.asarray(arr, dtype=np.float64)
    if a.ndim != 1:
        raise ValueError('1-D array expected')
    if a.size > 1 and a.strides[0] % 8:
        a = np.ascontiguousarray(a)
    dev = torch.device('cuda', torch.cuda.current_device() if device_index is None else int(device_index))
    out = torch.empty(a.size, dtype=torch.float64, device=dev)
    if a.size:
        _lib.check(_lib.load().ptg_upload_host_f64(a.size, a.ctypes.data, a.strides[0] // 8 if a.size > 1 else 1,
                                                   out.data_ptr(), dev.index))
    return out
