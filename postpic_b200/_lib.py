"""ctypes binding of libptg_b200.so -- the C ABI declared in include/ptg_b200.h.

There is deliberately NO fallback: if the shared library has not been built, or
no sm_100 GPU is present, the product path raises.  (The CPU oracle lives under
oracle/ and is test infrastructure; nothing here imports it.)
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PTG_B200_LIB") or os.path.join(_PKG, "lib", "libptg_b200.so")   # env: A/B builds

PTG_OK = 0
PTG_EINVAL = -1
PTG_ECUDA = -2
PTG_ENOMEM = -3
PTG_ENODEVICE = -4
PTG_EUNSUPPORTED = -5

ZERO_GRID = 0x1
METHOD_AUTO = 0x00
METHOD_GLOBAL = 0x10
METHOD_ROWS = 0x20
METHOD_TILED = 0x30
METHOD_SORTED = 0x40
METHOD_RUNS = 0x50
METHOD_RUNS_WIDE = 0x60
METHOD_FIXED = 0x70
EXACT_ACCUM = 0x100        # keep shared-memory tiles in fp64 (CAS loops) instead of 64-bit fixed point
METHODS = {None: METHOD_AUTO, 'auto': METHOD_AUTO, 'global': METHOD_GLOBAL, 'rows': METHOD_ROWS,
           'tiled': METHOD_TILED, 'sorted': METHOD_SORTED, 'runs': METHOD_RUNS, 'runs_wide': METHOD_RUNS_WIDE,
           'fixed': METHOD_FIXED, 'tiled_exact': METHOD_TILED | EXACT_ACCUM, 'auto_exact': METHOD_AUTO | EXACT_ACCUM}


class PtgError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "libptg_b200: %s (code %d)" % (message, code))
        self.code = code


class NoDeviceError(PtgError):
    pass


_c = ctypes
_P = _c.c_void_p
_I64 = _c.c_int64

class DepositArgs(_c.Structure):
    """ptg_deposit_args of include/ptg_b200.h (argument block of ptg_deposit_ex_f64)."""
    _fields_ = [("struct_size", _c.c_size_t), ("ndim", _c.c_int), ("shape", _c.c_int), ("n", _I64),
                ("x", _P), ("y", _P), ("z", _P), ("w", _P), ("strides", _P), ("lo", _P), ("inv", _P),
                ("bins", _P), ("grid", _P), ("flags", _c.c_uint), ("mask", _P), ("mask_stride", _I64),
                ("fuse_kind", _c.c_int), ("px", _P), ("py", _P), ("pz", _P), ("pstrides", _P),
                ("mass", _c.c_double), ("c", _c.c_double), ("elementary_charge", _c.c_double),
                ("workspace", _P), ("workspace_bytes", _c.c_size_t), ("stream", _P),
                ("axis_kind", _c.c_int * 3), ("pos_x", _P), ("pos_y", _P), ("pos_z", _P), ("pos_strides", _P),
                ("charge", _c.c_double), ("time", _c.c_double), ("atomic_mass", _c.c_double),
                ("weight_const", _c.c_double)]


# name -> (restype, argtypes); every symbol include/ptg_b200.h declares
SIGNATURES = {
    "ptg_abi_version": (_c.c_int, []),
    "ptg_last_error": (_c.c_char_p, []),
    "ptg_strerror": (_c.c_char_p, [_c.c_int]),
    "ptg_device_info": (_c.c_int, [_c.c_int, _c.c_char_p, _c.c_int, _P, _P, _P, _P, _P]),
    "ptg_axis_setup": (_c.c_int, [_c.c_double, _c.c_double, _c.c_int, _P, _P, _P]),
    "ptg_deposit_f64": (_c.c_int, [_c.c_int, _c.c_int, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                   _c.c_uint, _P, _c.c_size_t, _P]),
    "ptg_deposit_fused_f64": (_c.c_int, [_c.c_int, _c.c_int, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _c.c_uint,
                                         _c.c_int, _P, _P, _P, _P, _c.c_double, _c.c_double, _c.c_double, _P]),
    "ptg_deposit_ex_f64": (_c.c_int, [_P]),
    "ptg_scalar_f64": (_c.c_int, [_P, _c.c_int, _P]),
    "ptg_workspace_bytes": (_c.c_size_t, [_c.c_int, _c.c_int, _I64, _P, _c.c_uint]),
    "ptg_padded_cells": (_I64, [_c.c_int, _c.c_int, _P]),
    "ptg_minmax_f64": (_c.c_int, [_I64, _P, _I64, _P, _P]),
    "ptg_scale_f64": (_c.c_int, [_P, _I64, _c.c_double, _P]),
    "ptg_divide_f64": (_c.c_int, [_P, _I64, _c.c_double, _P]),
    "ptg_zero_f64": (_c.c_int, [_P, _I64, _P]),
    "ptg_weights_f64": (_c.c_int, [_c.c_int, _I64, _P, _P, _P, _P, _P, _c.c_double, _c.c_double, _c.c_double, _P, _P]),
    "ptg_histogram_host_f64": (_c.c_int, [_c.c_int, _c.c_int, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                          _c.c_uint, _c.c_int]),
    "ptg_deposit_host_f64": (_c.c_int, [_c.c_int, _c.c_int, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                        _c.c_uint, _c.c_int]),
    "ptg_minmax_host_f64": (_c.c_int, [_I64, _P, _I64, _P, _c.c_int]),
    "ptg_upload_host_f64": (_c.c_int, [_I64, _P, _I64, _P, _c.c_int]),
    "ptg_release": (_c.c_int, []),
    "ptg_selftest_div6": (_c.c_int, [_I64, _P, _P, _P]),
    "ptg_selftest_radix_sort": (_c.c_int, [_I64, _c.c_int, _P, _P, _P]),
    "ptg_debug_last_probe": (_c.c_int, [_P, _P, _P, _P, _P]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises ImportError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "postpic_b200: %s is missing. Build it with `python -m postpic_b200.build` "
                "(needs nvcc; there is no CPU fallback)." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError = ABI mismatch, on purpose
            fn.restype = res
            fn.argtypes = args
        if lib.ptg_abi_version() != 1:
            raise ImportError("postpic_b200: ABI version mismatch in %s" % LIB_PATH)
        _lib = lib
    return _lib


def check(rc):
    if rc == PTG_OK:
        return
    msg = load().ptg_last_error().decode("utf-8", "replace")
    if rc == PTG_EINVAL:
        raise ValueError(msg)
    if rc == PTG_ENODEVICE:
        raise NoDeviceError(rc, msg)
    if rc == PTG_ENOMEM:
        raise MemoryError(msg)
    raise PtgError(rc, msg)


def device_info(device=-1):
    lib = load()
    name = ctypes.create_string_buffer(256)
    sm = _c.c_int(0)
    l2 = _I64(0)
    hbm = _I64(0)
    major = _c.c_int(0)
    minor = _c.c_int(0)
    check(lib.ptg_device_info(device, name, 256, _c.byref(sm), _c.byref(l2), _c.byref(hbm),
                              _c.byref(major), _c.byref(minor)))
    return dict(name=name.value.decode(), sm_count=sm.value, l2_bytes=l2.value, hbm_bytes=hbm.value,
                cc=(major.value, minor.value))


PROBE_MODES = {0: 'rows', 1: 'tiled', 2: 'runs', 3: 'runs_wide'}


def last_probe():
    """Verdict of the most recent AUTO / TILED probe on the current device (test / diagnostics hook)."""
    lib = load()
    mode = _c.c_int(0)
    ext = (_c.c_int * 3)()
    hit, run, near = _c.c_float(0), _c.c_float(0), _c.c_float(0)
    check(lib.ptg_debug_last_probe(_c.byref(mode), ext, _c.byref(hit), _c.byref(run), _c.byref(near)))
    return dict(mode=PROBE_MODES.get(mode.value, mode.value), ext=list(ext), est_hit=hit.value,
                run_frac=run.value, near_frac=near.value)
