"""CPU-only tests: the C-ABI library loads and exports every symbol the header declares,
the host-side set-up arithmetic, argument normalisation of histogramdd and the
MultiSpecies dispatch (with the oracle standing in for the CUDA module -- allowed in
tests/ only), and the no-fallback guarantee."""
import ctypes
import os
import re

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from postpic_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "ptg_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ptg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.ptg_abi_version() == 1
    assert lib.ptg_strerror(-4) == b"no usable sm_100 GPU"


def test_library_contains_only_sm100a_code():
    import shutil
    import subprocess
    from postpic_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_axis_setup_matches_reference_arithmetic():
    from postpic_b200 import _lib
    from postpic_b200._setup import axis_setup
    lib = _lib.load()
    rng = np.random.default_rng(0)
    cases_ = [(0.0, 1.0, 10), (0.001, 0.999, 1000), (-1.0, 1.0, 1024), (2.0, 2.0, 4), (0.0, 0.0, 3),
              (1e-300, 1e-300, 5), (5.0, -5.0, 7), (0.0, 40e-6, 2048)]
    cases_ += [(float(a), float(a + abs(b)), int(n)) for a, b, n in
               zip(rng.normal(0, 10, 50), rng.normal(0, 3, 50), rng.integers(1, 5000, 50))]
    for lo, hi, bins in cases_:
        plo, pinv, edges = axis_setup((lo, hi), bins)
        clo, chi, cinv = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        assert lib.ptg_axis_setup(lo, hi, bins, ctypes.byref(clo), ctypes.byref(chi), ctypes.byref(cinv)) == 0
        assert clo.value == plo and cinv.value == pinv
        assert edges[0] == plo and len(edges) == bins + 1
        # the reference's expression, literally (pyx:53-60)
        xmin, xmax = lo, hi
        sx = np.spacing(xmax)
        if np.abs(xmax - xmin) < sx:
            xmax += sx
            xmin -= sx
        assert plo == xmin and np.array_equal(edges, np.linspace(xmin, xmax, bins + 1))
        with np.errstate(over='ignore', divide='ignore'):
            assert pinv == 1.0 / (xmax - xmin) * bins


def test_shape_spellings():
    from postpic_b200._setup import shape_order, shapes
    assert shapes == [[0, 'NGP'], [1, 'tophat'], [2, 'triangle'], [3, 'spline3']]
    for order, name in shapes:
        assert shape_order(order) == order and shape_order(name) == order
        assert shape_order(float(order)) == order and shape_order(np.int64(order)) == order
    assert shape_order(True) == 1 and shape_order(False) == 0
    for bad in (4, -1, 'cubic', None, 2.5):
        with pytest.raises(ValueError):
            shape_order(bad)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from postpic_b200 import _lib
    from postpic_b200.particles import _particlestogrid as ptg
    x = np.linspace(0.1, 0.9, 100)
    with pytest.raises(_lib.NoDeviceError):
        ptg.histogram(x, range=(0, 1), bins=10, shape=2)
    with pytest.raises(_lib.NoDeviceError):
        ptg.histogram2d(x, x, range=((0, 1), (0, 1)), bins=(4, 4), shape=0)
    with pytest.raises(_lib.NoDeviceError):
        _lib.device_info()
    # the front door too, also on the large-array / range-from-data route (device min/max + upload)
    from postpic_b200 import histogramdd
    big = np.random.default_rng(0).random(1 << 21)
    with pytest.raises(_lib.NoDeviceError):
        histogramdd((big, big), bins=(8, 8))
    with pytest.raises(_lib.NoDeviceError):
        histogramdd(x, bins=8, range=(0, 1))
    # argument errors are still reported before any device work
    with pytest.raises(ValueError):
        ptg.histogram2d(x, x[:5], range=((0, 1), (0, 1)), bins=(4, 4))
    with pytest.raises(ValueError):
        ptg.histogram(x, range=(0, 1), bins=4, shape=7)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "postpic_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "ptg_oracle" not in text, f


@pytest.fixture
def cpu_stack(monkeypatch, oracle):
    """The Python host stack with the oracle in place of the CUDA module."""
    from postpic_b200.particles import _routines
    monkeypatch.setattr(_routines, "ptg", oracle)
    return _routines


def test_histogramdd_argument_forms(cpu_stack, oracle):
    hdd = cpu_stack.histogramdd
    rng = np.random.default_rng(3)
    x, y, z, w = rng.random(1000), 2 * rng.random(1000), 3 * rng.random(1000), rng.random(1000)
    h0, (e0,) = hdd(x, bins=20, range=(0, 1), shape=0)
    for h, (e,) in [hdd((x,), bins=20, range=((0, 1),), shape=0), hdd((x,), bins=20, range=(0, 1), shape=0),
                    hdd(x, bins=20, range=((0, 1),), shape=0), hdd(x, bins=(20,), range=((0, 1),), shape=0)]:
        assert np.array_equal(h, h0) and np.array_equal(e, e0)
    assert np.array_equal(h0, np.histogram(x, bins=20, range=(0, 1))[0])
    # (N, D) array form == tuple form; default shape is 2; shape=None means default
    nd = np.asarray([x, y]).T
    a, ea = hdd(nd, bins=[20, 22], range=((0, 1), (0, 2)), weights=w)
    b, eb = hdd((x, y), bins=[20, 22], range=((0, 1), (0, 2)), weights=w, shape=None)
    c = oracle.histogram2d(x, y, weights=w, range=((0, 1), (0, 2)), bins=(20, 22), shape=2)[0]
    assert np.array_equal(a, b) and np.array_equal(a, c)
    # default bins, scalar bins, range completion from the data
    assert hdd(x, range=(0, 1))[0].shape == (800,)
    assert hdd((x, y), range=((0, 1), (0, 2)))[0].shape == (500, 500)
    assert hdd((x, y, z), bins=27, range=((0, 1), (0, 2), (0, 3)), shape=0)[0].shape == (27, 27, 27)
    h, (ex, ey) = hdd((x, y), bins=(8, 8), range=((None, 0.9), None))
    assert ex[0] == x.min() and ex[-1] == 0.9 and ey[0] == y.min() and ey[-1] == y.max()
    # lists are converted, ints are cast
    assert np.array_equal(hdd(list(x), bins=20, range=(0, 1), shape=0)[0], h0)
    assert hdd(np.arange(10), bins=5, range=(0, 10), shape=0)[0].sum() == 9
    with pytest.raises(TypeError):
        hdd(x, range=(0, 1), foo=3)
    with pytest.raises(ValueError):
        hdd((x, x, x, x), bins=3)
    with pytest.raises(ValueError):
        hdd((x, y), bins=4, range=(((0, 1), 1), (0, 1)))     # malformed range: non-scalar end


def test_multispecies_dispatch_on_cpu_stack(cpu_stack, monkeypatch, oracle):
    from postpic_b200.particles import particles as P
    monkeypatch.setattr(P, "histogramdd", cpu_stack.histogramdd)
    rng = np.random.default_rng(0)
    n = 5000
    me, c = 9.109383e-31, 299792458.0
    src = P.ArraySource('electron', dict(x=rng.normal(0, 1, n), y=rng.normal(0, 1, n), z=rng.normal(0, 1, n),
                                         px=rng.normal(0, 1, n) * me * c, py=rng.normal(0, 1, n) * me * c,
                                         pz=rng.normal(0, 1, n) * me * c, weight=2.0, mass=me),
                        extent={'x': (-4, 4), 'y': (-4, 4)})
    p = P.MultiSpecies(src)
    f1 = p.createField('x', 'y', bins=[50, 40], simextent=True)
    f2 = (p + p).createField('x', 'y', bins=[50, 40], simextent=True)
    assert np.allclose(2 * f1.matrix, f2.matrix)                    # test/test_particles.py:37-49
    # several host species are deposited one by one into one accumulator instead of being concatenated:
    # same result as the deposit of the concatenated arrays, ghost cells included, with and without ranges
    other = P.ArraySource('ion', dict(x=rng.normal(1, 2, 3000), y=rng.normal(0, 1, 3000), z=np.zeros(3000),
                                      px=np.zeros(3000), py=np.zeros(3000), pz=np.zeros(3000), weight=0.5,
                                      mass=1836 * me))
    both = P.MultiSpecies(src, other)
    xs = np.hstack([src.atomic('x'), other.atomic('x')])
    ys = np.hstack([src.atomic('y'), other.atomic('y')])
    ws = np.hstack([np.full(n, 2.0), np.full(3000, 0.5)])
    for kw, rngs in ((dict(rangex=(-4, 4), rangey=(-3, 3)), ((-4, 4), (-3, 3))),
                     (dict(), ((xs.min(), xs.max()), (ys.min(), ys.max())))):
        hb, eb = both._createHistgram('x', 'y', bins=[30, 20], shape=3, **kw)
        hc, ex, ey = oracle.histogram2d(xs, ys, weights=ws, range=rngs, bins=(30, 20), shape=3)
        assert np.allclose(hb, hc / ((ex[1] - ex[0]) * (ey[1] - ey[0])), rtol=1e-12, atol=1e-300)
        assert hb.base is not None and hb.base.shape == hc.base.shape
        assert np.array_equal(eb[0], ex) and np.array_equal(eb[1], ey)
    fb = both.filter('x > 0').createField('x', bins=[16], rangex=(-4, 4), shape=1)
    keep = xs > 0
    h1, e1 = oracle.histogram(xs[keep], weights=ws[keep], range=(-4, 4), bins=16, shape=1)
    assert np.allclose(fb.matrix, h1 / (e1[1] - e1[0]), rtol=1e-12) and fb.infostring == '%d npart in 2 species' % keep.sum()
    # 1/dV normalisation, weight * (expr), default shape 2
    h, ex, ey = oracle.histogram2d(src.atomic('x'), src.atomic('y'), weights=np.full(n, 2.0),
                                   range=((-4, 4), (-4, 4)), bins=(50, 40), shape=2)
    assert np.allclose(f1.matrix, h / ((ex[1] - ex[0]) * (ey[1] - ey[0])), rtol=1e-14)
    assert f1.axes[0].name == 'x' and f1.axes[1].unit == 'm' and f1.infostring == '5000 npart in 1 species'
    # energy weights
    import scipy.constants as sc
    f3 = p.createField('x', weights='Ekin_MeV', bins=[32], rangex=(-3, 3), shape=3)
    a = src._arrays
    np2 = (a['px'] ** 2 + a['py'] ** 2 + a['pz'] ** 2) / (me * sc.c) ** 2
    ek = 2.0 * (np2 / (np.sqrt(1 + np2) + 1) * me * sc.c ** 2 / sc.elementary_charge / 1e6)
    h, e = oracle.histogram(a['x'], weights=ek, range=(-3, 3), bins=32, shape=3)
    assert np.allclose(f3.matrix, h / (e[1] - e[0]), rtol=1e-12)
    assert f3.name == 'Ekin_MeVelectron'
    # filter(): host species are compressed, like MultiSpecies.filter in the reference (particles.py:561-582)
    sel = p.filter('x > 0').filter('gamma > 1.1')
    keep = (a['x'] > 0) & (np2 / (np.sqrt(1 + np2) + 1) + 1 > 1.1)
    assert sel.npart == int(keep.sum()) and 0 < sel.npart < n
    f4 = sel.createField('x', 'y', bins=[16, 12], rangex=(-3, 3), rangey=(-3, 3), shape=1)
    h, ex, ey = oracle.histogram2d(a['x'][keep], a['y'][keep], weights=np.full(int(keep.sum()), 2.0),
                                   range=((-3, 3), (-3, 3)), bins=(16, 12), shape=1)
    assert np.allclose(f4.matrix, h / ((ex[1] - ex[0]) * (ey[1] - ey[0])), rtol=1e-14)
    assert f4.infostring == '%d npart in 1 species' % keep.sum()
    # simgrid is a no-op for bins (the reference's simgridpoints returns None)
    assert p.createField('x', 'y', bins=[7, 9], simgrid=True).matrix.shape == (7, 9)
    with pytest.raises(TypeError):
        p.createField('x', bogus=1)
    with pytest.raises(KeyError):
        p.createField('nonexistent_scalar', bins=[4], rangex=(0, 1))
    h, edges = p._createHistgram('nonexistent_scalar', bins=[4], rangex=(0, 1), force=True)
    assert h.shape == (4,) and not h.any()


def test_field_container_contract():
    """`Field(matrix, xedges=..., ...)` (datahandling.py:561-612): one edge array per dimension, host matrix
    kept as numpy; the lazily materialised device grid of SURVEY section 8 row f4 behaves like a host one."""
    from postpic_b200.particles.particles import Field, _DeviceGrid
    m = np.arange(12.0).reshape(3, 4)
    f = Field(m, xedges=np.linspace(0, 1, 4), yedges=np.linspace(-1, 1, 5))
    assert f.shape == (3, 4) and f.dimensions == 2 and f.device_matrix is None
    assert f.extent == [0.0, 1.0, -1.0, 1.0] and np.allclose(f.axes[1].grid, [-0.75, -0.25, 0.25, 0.75])
    assert np.asarray(f) is f.matrix
    with pytest.raises(ValueError):
        Field(m, xedges=np.linspace(0, 1, 4))

    class FakeTensor(object):                      # stands in for a CUDA tensor: slicing and a host copy
        def __init__(self, a):
            self.a, self.copies = a, 0
            self.shape = a.shape

        def __getitem__(self, idx):
            return FakeTensor(self.a[idx])

        def cpu(self):
            self.copies += 1
            return self

        def numpy(self):
            return self.a

    padded = FakeTensor(np.pad(m, 2))
    g = Field(_DeviceGrid(padded, (slice(2, 5), slice(2, 6))), xedges=np.linspace(0, 1, 4), yedges=np.linspace(-1, 1, 5))
    assert g.shape == (3, 4) and g.dimensions == 2 and padded.copies == 0          # nothing copied yet
    assert np.array_equal(g.matrix, m) and padded.copies == 1
    assert np.array_equal(np.asarray(g), m) and padded.copies == 1                   # cached
    g.matrix = 2 * m                                                                 # h-like in-place users replace it
    assert g.device_matrix is None and np.array_equal(g.matrix, 2 * m)


def test_dummy_species_follows_the_reference_recipe():
    # datareader/dummy.py:43-61 restated: seeded, independent of the caller's RNG state
    from postpic_b200 import ArraySource
    np.random.seed(5)
    before = np.random.random()
    np.random.seed(5)
    src = ArraySource.dummy(10000, dimensions=3, seed=0)
    assert np.random.random() == before                     # the global RNG is left untouched
    np.random.seed(0)
    x, y, z = (np.random.normal(size=10000) for _ in range(3))
    mc = 9.109383e-31 * 299792458.0
    assert np.array_equal(src.atomic('x'), x) and np.array_equal(src.atomic('y'), y) and np.array_equal(src.atomic('z'), z)
    assert np.array_equal(src.atomic('px'), np.roll(x, 1) ** 2 * mc) and np.array_equal(src.atomic('pz'), np.roll(x, 3) * mc)
    assert np.array_equal(src.atomic('weight'), np.ones(10000)) and len(src) == 10000
    assert sorted(src.atomic('id')) == list(range(10000))
    assert src.extent == {'x': (-2.0, 10.0), 'y': (-5.0, 5.0), 'z': (-4.0, 4.0)} and src.species == 'electron'
    assert set(ArraySource.dummy(10, dimensions=1).extent) == {'x'}


def test_shard_bounds_cover_everything():
    from postpic_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 1000003):
        for ws in (1, 2, 3, 8):
            spans = [shard_bounds(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_expression_evaluator_accepts_arithmetic_only():
    """The scalar-expression evaluator (the reference uses numexpr, particles.py:290-330): arithmetic, comparisons and
    calls of the known functions; everything that could reach Python internals is rejected before evaluation."""
    import numpy as np
    import pytest
    from postpic_b200 import MultiSpecies, ArraySource
    ms = MultiSpecies(ArraySource.dummy(500, 2, seed=1))
    x, y = np.asarray(ms('x')), np.asarray(ms('y'))
    assert np.allclose(ms('x*2 + sqrt(abs(y)) - 3'), x * 2 + np.sqrt(np.abs(y)) - 3)
    assert np.array_equal(ms('(x > 0) & (y < 0.5)'), (x > 0) & (y < 0.5))
    assert np.allclose(ms('arctan2(y, x) ** 2'), np.arctan2(y, x) ** 2)
    for bad in ("().__class__", "x.real", "[v for v in (1,)]", "__import__('os')", "lambda: 1", "x[0]",
                "sqrt(x, out=y)", "'a' + 'b'", "x if y else x"):
        with pytest.raises(ValueError):
            ms(bad)
    with pytest.raises(KeyError, match='"nope" not found!'):
        ms('nope + 1')


def test_numa_helpers_do_not_raise():
    """parallel.bind_to_gpu_numa's helpers are best effort: on a box without NVML / NUMA information they return
    None / False instead of raising, and numa_pages_of reads this process's own map."""
    import numpy as np
    from postpic_b200 import parallel
    assert parallel.gpu_numa_node(0) is None or parallel.gpu_numa_node(0) >= 0
    a = np.ones(1 << 20)
    pages = parallel.numa_pages_of(a.ctypes.data)
    assert pages is None or (isinstance(pages, dict) and all(v >= 0 for v in pages.values()))
    assert parallel.prefer_numa_node(0) in (True, False)
    cpus = parallel.bind_to_gpu_numa(0)
    assert cpus is None or len(cpus) > 0
