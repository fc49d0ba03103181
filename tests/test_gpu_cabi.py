"""The C ABI called directly through ctypes (no Python drop-in in between): return codes,
error messages, asynchronous device entry on a caller stream, concurrent streams, 64-bit
particle counts."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from postpic_b200 import _lib
    return _lib.load()


def _axis(lo, hi, bins):
    return lo, 1.0 / (hi - lo) * bins


def test_error_codes_and_messages(lib):
    import torch
    from postpic_b200 import _lib
    x = torch.rand(100, dtype=torch.float64, device='cuda')
    grid = torch.zeros(14, dtype=torch.float64, device='cuda')
    lo = np.zeros(3)
    inv = np.full(3, 10.0)
    bins = (ctypes.c_int * 3)(10, 1, 1)

    def call(ndim=1, shape=2, n=100, xp=None, g=None, b=bins, flags=_lib.ZERO_GRID | _lib.METHOD_ROWS):
        return lib.ptg_deposit_f64(ndim, shape, n, x.data_ptr() if xp is None else xp, None, None, None, None,
                                   lo.ctypes.data, inv.ctypes.data, b, grid.data_ptr() if g is None else g,
                                   flags, None, 0, None)
    assert call() == _lib.PTG_OK
    torch.cuda.synchronize()
    assert abs(float(grid.sum()) - 100.0) < 1e-9
    assert call(ndim=4) == _lib.PTG_EINVAL and b"ndim" in lib.ptg_last_error()
    assert call(shape=7) == _lib.PTG_EINVAL and b"shape" in lib.ptg_last_error()
    assert call(n=-1) == _lib.PTG_EINVAL
    assert call(ndim=2) == _lib.PTG_EINVAL                      # y pointer missing
    assert call(b=(ctypes.c_int * 3)(0, 1, 1)) == _lib.PTG_EINVAL
    assert call(flags=0xF0) == _lib.PTG_EUNSUPPORTED
    huge = (ctypes.c_int * 3)(2 ** 31 - 1, 1, 1)
    assert call(b=huge) == _lib.PTG_EUNSUPPORTED and b"2^31" in lib.ptg_last_error()
    assert lib.ptg_strerror(_lib.PTG_EUNSUPPORTED) == b"unsupported request"
    assert lib.ptg_padded_cells(3, 1, (ctypes.c_int * 3)(256, 256, 256)) == 258 ** 3
    assert lib.ptg_workspace_bytes(2, 2, 1000, (ctypes.c_int * 3)(10, 10, 1), _lib.METHOD_ROWS) == 0
    assert lib.ptg_workspace_bytes(2, 2, 1000, (ctypes.c_int * 3)(10, 10, 1), _lib.METHOD_SORTED) > 16000
    # a caller-supplied workspace that is too small is refused, a big enough one is used
    ws = torch.empty(1 << 20, dtype=torch.uint8, device='cuda')
    rc = lib.ptg_deposit_f64(1, 2, 100, x.data_ptr(), None, None, None, None, lo.ctypes.data, inv.ctypes.data, bins,
                             grid.data_ptr(), _lib.ZERO_GRID | _lib.METHOD_SORTED, ws.data_ptr(), 1024, None)
    assert rc == _lib.PTG_ENOMEM
    need = lib.ptg_workspace_bytes(1, 2, 100, bins, _lib.METHOD_SORTED)
    ws = torch.empty(need, dtype=torch.uint8, device='cuda')
    rc = lib.ptg_deposit_f64(1, 2, 100, x.data_ptr(), None, None, None, None, lo.ctypes.data, inv.ctypes.data, bins,
                             grid.data_ptr(), _lib.ZERO_GRID | _lib.METHOD_SORTED, ws.data_ptr(), need, None)
    assert rc == _lib.PTG_OK
    torch.cuda.synchronize()
    assert abs(float(grid.sum()) - 100.0) < 1e-9


def test_host_entry_matches_device_entry_and_accumulate_flag(lib, oracle):
    import torch
    from postpic_b200 import _lib
    rng = np.random.default_rng(4)
    n = 300000
    x, y, w = rng.uniform(0, 1, n), rng.uniform(0, 2, n), rng.uniform(0, 1, n)
    lo = np.array([0.0, 0.0, 0.0])
    inv = np.array([_axis(0, 1, 40)[1], _axis(0, 2, 30)[1], 1.0])
    bins = (ctypes.c_int * 3)(40, 30, 1)
    host = np.empty((44, 34))
    assert lib.ptg_histogram_host_f64(2, 2, n, x.ctypes.data, y.ctypes.data, None, w.ctypes.data, None,
                                      lo.ctypes.data, inv.ctypes.data, bins, host.ctypes.data, 0, -1) == 0
    ref = oracle.histogram2d(x, y, weights=w, range=((0, 1), (0, 2)), bins=(40, 30), shape=2)[0].base
    assert np.max(np.abs(host - ref)) <= 1e-12 * np.max(np.abs(ref))
    # strided host arrays through the C entry (2-D memcpy path): every 2nd particle
    strides = (ctypes.c_int64 * 4)(2, 2, 1, 2)
    assert lib.ptg_histogram_host_f64(2, 2, n // 2, x.ctypes.data, y.ctypes.data, None, w.ctypes.data, strides,
                                      lo.ctypes.data, inv.ctypes.data, bins, host.ctypes.data, 0, -1) == 0
    ref2 = oracle.histogram2d(x[::2], y[::2], weights=w[::2], range=((0, 1), (0, 2)), bins=(40, 30), shape=2)[0].base
    assert np.max(np.abs(host - ref2)) <= 1e-12 * np.max(np.abs(ref2))
    # device entry without PTG_ZERO_GRID accumulates onto what is there: two halves == whole
    tx, ty, tw = (torch.from_numpy(a).cuda() for a in (x, y, w))
    grid = torch.zeros((44, 34), dtype=torch.float64, device='cuda')
    h = n // 2
    for sl in (slice(0, h), slice(h, n)):
        assert lib.ptg_deposit_f64(2, 2, sl.stop - sl.start, tx[sl].data_ptr(), ty[sl].data_ptr(), None,
                                   tw[sl].data_ptr(), None, lo.ctypes.data, inv.ctypes.data, bins,
                                   grid.data_ptr(), _lib.METHOD_AUTO, None, 0,
                                   torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    assert np.max(np.abs(grid.cpu().numpy() - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_concurrent_streams_into_one_grid(lib):
    """Four streams deposit four quarters of a cloud into the same grid at the same time (what the
    host pipeline does internally); accumulation is by reductions, so the sum must be complete."""
    import torch
    from postpic_b200 import device
    n = 8_000_000
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    y = torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    w = torch.randint(1, 4, (n,), device='cuda', generator=g).to(torch.float64)
    for shape, method in ((0, None), (2, 'rows'), (2, None), (1, 'tiled')):
        plan = device.Plan([(0, 1), (0, 1)], [512, 512], shape)
        grid = torch.zeros(plan.dims, dtype=torch.float64, device='cuda')
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream() for _ in range(4)]
        q = n // 4
        for k, st in enumerate(streams):
            sl = slice(k * q, (k + 1) * q)
            device.deposit_into(grid, plan, [x[sl], y[sl]], w[sl], method=method, zero=False, stream=st.cuda_stream)
        torch.cuda.synchronize()
        whole, _ = device.deposit([x, y], w, range=[(0, 1), (0, 1)], bins=(512, 512), shape=shape, method=method)
        if shape == 0:
            assert torch.equal(grid, whole)
        else:
            assert float((grid - whole).abs().max()) <= 1e-12 * float(whole.abs().max())


def test_host_entries_from_several_threads(oracle):
    """The host-buffer entry points share one staging ring per device and serialise on a per-device
    mutex; ctypes releases the GIL, so Python threads really do call them concurrently.  Every thread
    must get the histogram of ITS data (different shapes, sizes and dimensionalities per thread)."""
    import threading
    import postpic_b200.particles._particlestogrid as ptg
    from postpic_b200 import histogramdd
    rng = np.random.default_rng(21)
    jobs = []
    for t in range(6):
        n = 200000 + 50000 * t
        shape = t % 4
        if t % 2:
            data = (rng.uniform(0, 1, n), rng.uniform(-1, 1, n))
            kw = dict(range=((0, 1), (-1, 1)), bins=(64 + t, 48), shape=shape, weights=rng.uniform(0, 1, n))
            want = oracle.histogram2d(data[0], data[1], **kw)[0]
        else:
            data = (rng.normal(0, 1, n),)
            kw = dict(range=(-3, 3), bins=200 + t, shape=shape, weights=rng.uniform(0, 1, n))
            want = oracle.histogram(data[0], **kw)[0]
        jobs.append((data, kw, want))
    big = rng.normal(0, 1, (1 << 20) + 777)           # large enough for the device min/max + upload route
    got = [None] * len(jobs)
    errors = []

    def run(k):
        try:
            for _ in range(3):
                data, kw, _ = jobs[k]
                got[k] = (ptg.histogram2d(data[0], data[1], **kw) if len(data) == 2 else ptg.histogram(data[0], **kw))[0]
            # the range-less route (min/max + upload through the same ring) from a thread as well
            if len(jobs[k][0]) == 1:
                h, _ = histogramdd(big, bins=50, shape=0)
                assert h.sum() == big.size - 2                    # strict edges drop the min and the max
        except Exception as e:                                    # pragma: no cover
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=run, args=(k,)) for k in range(len(jobs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=120)
    assert not errors, errors
    for (data, kw, want), h in zip(jobs, got):
        assert h is not None and h.shape == want.shape
        assert np.max(np.abs(h - want)) <= 1e-12 * np.max(np.abs(want))


def test_more_than_2_31_particles(lib):
    """The reference's `cdef int n` overflows at 2^31 particles; here indices are 64-bit.
    1-D, unweighted, shape 0: every particle must be counted exactly once (17 GB of input)."""
    import torch
    from postpic_b200 import device
    free, _ = torch.cuda.mem_get_info()
    n = 2 ** 31 + 4097
    if free < n * 8 + (2 << 30):
        pytest.skip("not enough free HBM for a 2^31-particle array")
    x = torch.empty(n, dtype=torch.float64, device='cuda')
    x.uniform_(0.0001, 0.9999)
    for method in ('global', None, 'runs'):
        grid, plan = device.deposit([x], None, range=(0, 1), bins=1000, shape=0, method=method)
        assert float(grid.sum()) == float(n), method
    grid, plan = device.deposit([x], None, range=(0, 1), bins=1000, shape=1, method=None)
    assert abs(float(grid.sum()) - n) <= 1e-12 * n
    del x


@pytest.mark.parametrize("n,bits,maxkey", [(1, 8, 200), (33, 8, 7), (8192, 16, 65535), (8193, 21, 2 ** 21 - 1),
                                           (1_000_003, 24, 2 ** 21), (5_000_000, 31, 2 ** 31 - 1),
                                           (3_000_000, 8, 3)])
def test_radix_sort_is_a_stable_sort(lib, n, bits, maxkey):
    """The hand-written LSD radix sort of method SORTED against torch's stable sort: same sorted keys,
    same permutation (equal keys keep their input order)."""
    import torch
    from postpic_b200 import _lib
    g = torch.Generator(device='cuda').manual_seed(n + bits)
    keys = torch.randint(0, maxkey + 1, (n,), device='cuda', generator=g, dtype=torch.int64)
    ref_keys, ref_perm = torch.sort(keys, stable=True)
    k32 = keys.to(torch.int32).contiguous()           # bit pattern of the uint32 keys (all < 2^31 here)
    perm = torch.empty(n, dtype=torch.int32, device='cuda')
    _lib.check(lib.ptg_selftest_radix_sort(n, bits, k32.data_ptr(), perm.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(k32.to(torch.int64), ref_keys)
    assert torch.equal(perm.to(torch.int64), ref_perm)
