"""The reference's own hot-path tests (test/test_particlestogrid.py, test/test_particles.py:37-49)
re-hosted on the CUDA drop-in: same calls, same assertions, seeded data."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cf():
    from postpic_b200.particles import _particlestogrid
    _particlestogrid.method = None
    return _particlestogrid


@pytest.fixture(scope="module")
def histogramdd():
    from postpic_b200.particles import histogramdd
    return histogramdd


class TestHistogram:
    @pytest.fixture(autouse=True)
    def _data(self):
        rng = np.random.default_rng(11)
        self.data = rng.random(int(1e5))
        self.weights = rng.random(int(1e5))

    def test_histogram(self, cf):
        cfh, cfe = cf.histogram(self.data, bins=20, range=(0, 1), shape=0)
        nph, npe = np.histogram(self.data, bins=20, range=(0, 1))
        assert list(nph) == list(cfh)
        assert list(npe) == list(cfe)

    def test_histogramw(self, cf):
        cfh, cfe = cf.histogram(self.data, bins=100, range=(0, 1), shape=0, weights=self.weights)
        nph, npe = np.histogram(self.data, bins=100, range=(0, 1), weights=self.weights)
        assert abs(np.sum(np.abs(cfh - nph))) < 5e-8
        assert list(npe) == list(cfe)

    @pytest.mark.parametrize("shape", [1, 2, 3])
    @pytest.mark.parametrize("weighted", [False, True])
    def test_histogram_orders(self, cf, shape, weighted):
        w = self.weights if weighted else None
        cfh, cfe = cf.histogram(self.data, bins=100, range=(0.0, 1.0), shape=shape, weights=w)
        nph, npe = np.histogram(self.data, bins=100, range=(0.0, 1.0), weights=w)
        totalmass = np.sum(w) if weighted else len(self.data)
        assert abs(np.sum(cfh.base) - totalmass) < 5e-8      # cfh.base includes the ghost cells
        assert list(npe) == list(cfe)


class TestHistogram2d3d:
    @pytest.fixture(autouse=True)
    def _data(self):
        rng = np.random.default_rng(12)
        self.datax = rng.random(1000)
        self.datay = 2 * rng.random(1000)
        self.dataz = 3 * rng.random(1000)
        self.weights = rng.random(1000)

    @pytest.mark.parametrize("shape", [0, 1, 2, 3])
    @pytest.mark.parametrize("weighted", [False, True])
    def test_histogram2d(self, cf, shape, weighted):
        w = self.weights if weighted else None
        cfh, cfex, cfey = cf.histogram2d(self.datax, self.datay, weights=w, bins=(20, 30),
                                         range=((0, 1), (0, 2)), shape=shape)
        nph, npex, npey = np.histogram2d(self.datax, self.datay, weights=w, bins=(20, 30), range=((0, 1), (0, 2)))
        totalmass = np.sum(w) if weighted else len(self.datax)
        assert abs(np.sum(cfh.base) - totalmass) < 5e-8
        assert list(npex) == list(cfex) and list(npey) == list(cfey)
        if shape == 0:
            assert np.allclose(cfh, nph)

    @pytest.mark.parametrize("shape", [0, 1, 2, 3])
    @pytest.mark.parametrize("weighted", [False, True])
    def test_histogram3d(self, cf, shape, weighted):
        w = self.weights if weighted else None
        cfh, ex, ey, ez = cf.histogram3d(self.datax, self.datay, self.dataz, weights=w, bins=(20, 30, 40),
                                         range=((0, 1), (0, 2), (0, 3)), shape=shape)
        totalmass = np.sum(w) if weighted else len(self.datax)
        assert abs(np.sum(cfh.base) - totalmass) < 5e-8


class TestHistogramdd:
    @pytest.fixture(autouse=True)
    def _data(self):
        rng = np.random.default_rng(13)
        self.datax = rng.random(1000)
        self.datay = 2 * rng.random(1000)
        self.dataz = 3 * rng.random(1000)
        self.weights = rng.random(1000)

    def test_histogram1d_call_forms(self, histogramdd):
        h0, (ex0,) = histogramdd(self.datax, bins=20, range=(0, 1), shape=0)
        forms = [histogramdd((self.datax,), bins=20, range=((0, 1),), shape=0),
                 histogramdd((self.datax,), bins=20, range=(0, 1), shape=0),
                 histogramdd(self.datax, bins=20, range=((0, 1),), shape=0),
                 histogramdd(self.datax, bins=(20,), range=((0, 1),), shape=0)]
        for h, (ex,) in forms:
            assert np.all(np.isclose(h0, h)) and np.all(np.isclose(ex0, ex))

    def test_against_numpy_1d(self, histogramdd):
        h0, (ex0,) = histogramdd(self.datax, bins=20, range=(0, 1), shape=0)
        h1, (ex1,) = np.histogramdd(self.datax, bins=20, range=((0, 1),))
        assert np.all(np.isclose(h0, h1)) and np.all(np.isclose(ex0, ex1))
        assert abs(np.sum(h0) - len(self.datax)) < 1e-7

    @pytest.mark.parametrize("weighted", [False, True])
    def test_against_numpy_2d_3d(self, histogramdd, weighted):
        w = self.weights if weighted else None
        total = np.sum(w) if weighted else len(self.datax)
        arg = (self.datax, self.datay)
        h0, e0 = histogramdd(arg, bins=[20, 22], range=((0, 1), (0, 2)), shape=0, weights=w)
        h1, e1 = np.histogramdd(arg, bins=[20, 22], range=((0, 1), (0, 2)), weights=w)
        assert np.all(np.isclose(h0, h1)) and all(np.all(np.isclose(a, b)) for a, b in zip(e0, e1))
        assert abs(np.sum(h0) - total) < 1e-7
        arg = (self.datax, self.datay, self.dataz)
        h0, e0 = histogramdd(arg, bins=[20, 22, 25], range=((0, 1), (0, 2), (0, 3)), shape=0, weights=w)
        h1, e1 = np.histogramdd(arg, bins=[20, 22, 25], range=((0, 1), (0, 2), (0, 3)), weights=w)
        assert np.all(np.isclose(h0, h1)) and all(np.all(np.isclose(a, b)) for a, b in zip(e0, e1))
        assert abs(np.sum(h0) - total) < 1e-7

    def test_nd_array_form_and_scalar_bins(self, histogramdd):
        arg = np.asarray([self.datax, self.datay]).T
        kw = dict(bins=[20, 22], range=((0, 1), (0, 2)), weights=self.weights)
        h0, e0 = histogramdd(arg, shape=0, **kw)
        h1, e1 = np.histogramdd(arg, **kw)
        assert np.all(np.isclose(h0, h1))
        arg = (self.datax, self.datay, self.dataz)
        h0, e0 = histogramdd(arg, bins=27, range=((0, 1), (0, 2), (0, 3)), shape=0, weights=self.weights)
        h1, e1 = np.histogramdd(arg, bins=27, range=((0, 1), (0, 2), (0, 3)), weights=self.weights)
        assert np.all(np.isclose(h0, h1)) and list(h0.shape) == [27, 27, 27]

    def test_range_from_data_host_and_device(self, histogramdd):
        import torch
        h0, e0 = histogramdd((self.datax, self.datay), bins=(16, 16), shape=2, weights=self.weights)
        assert e0[0][0] == self.datax.min() and e0[0][-1] == self.datax.max()
        t = [torch.from_numpy(a).cuda() for a in (self.datax, self.datay)]
        h1, e1 = histogramdd(t, bins=(16, 16), shape=2, weights=torch.from_numpy(self.weights).cuda())
        assert all(np.array_equal(a, b) for a, b in zip(e0, e1))
        assert np.allclose(h0, h1, rtol=1e-12, atol=1e-15)
        h2, e2 = histogramdd((self.datax, self.datay), bins=(16, 16), range=((None, 0.9), None), shape=2)
        assert e2[0][0] == self.datax.min() and e2[0][-1] == 0.9 and e2[1][-1] == self.datay.max()

    def test_range_from_large_host_arrays_uses_the_device_reduction(self, histogramdd):
        rng = np.random.default_rng(3)
        n = (1 << 21) + 12345
        x, y = rng.normal(0, 1, n), rng.normal(5, 2, n)
        h, (ex, ey) = histogramdd((x, y[::-1]), bins=(40, 30), shape=1)
        assert ex[0] == x.min() and ex[-1] == x.max() and ey[0] == y.min() and ey[-1] == y.max()
        assert abs(h.base.sum() - n) < 1e-6 * n
        x[77] = np.nan                                       # NaN propagates into the edges, like numpy
        _, (ex, _) = histogramdd((x, y), bins=(4, 4), shape=0)
        assert np.isnan(ex).all()

    def test_defaults_and_errors(self, histogramdd):
        h, (e,) = histogramdd(self.datax, range=(0, 1))
        assert h.shape == (800,)                          # default bins, default shape 2
        h2, _ = histogramdd(self.datax, range=(0, 1), shape=None)
        assert np.allclose(h, h2, rtol=1e-13, atol=0)
        with pytest.raises(TypeError):
            histogramdd(self.datax, range=(0, 1), nonsense=1)
        with pytest.raises(ValueError):
            histogramdd((self.datax,) * 4, bins=4)


class TestCreateField:
    """test/test_particles.py:37-49 -- createField of p + p is twice createField of p."""

    def _species(self, n=10000, device=None):
        from postpic_b200 import ArraySource
        rng = np.random.default_rng(0)
        me, c = 9.109383e-31, 299792458.0
        arrs = dict(x=rng.normal(0, 1, n), y=rng.normal(0, 1, n), z=rng.normal(0, 1, n),
                    px=rng.normal(0, 1, n) ** 2 * me * c, py=rng.normal(0, 1, n) * me * c,
                    pz=rng.normal(0, 1, n) * me * c, weight=1.0, mass=me, charge=-1.602176565e-19)
        if device is not None:
            import torch
            arrs = {k: (torch.from_numpy(v).to(device) if isinstance(v, np.ndarray) else v) for k, v in arrs.items()}
        return ArraySource('electron', arrs, extent={'x': (-4, 4), 'y': (-4, 4), 'z': (-4, 4)})

    def test_linearity(self):
        from postpic_b200 import MultiSpecies
        p = MultiSpecies(self._species())
        pp = p + p
        f1 = p.createField('x', 'y', bins=[50, 40], simextent=True)
        f2 = pp.createField('x', 'y', bins=[50, 40], simextent=True)
        assert np.allclose(2 * f1.matrix, f2.matrix)
        assert f1.matrix.shape == (50, 40) and f1.axes[0].name == 'x' and f1.axes[0].unit == 'm'

    def test_energy_weighted_matches_numpy_formula_and_device(self, cf):
        from postpic_b200 import MultiSpecies
        src = self._species()
        p = MultiSpecies(src)
        kw = dict(weights='Ekin_MeV', bins=[64, 64], shape=3, rangex=(-3, 3), rangey=(-3, 3))
        f = p.createField('x', 'y', **kw)
        # independent evaluation of weight * Ekin_MeV and of the 1/dV normalisation
        import scipy.constants as sc
        a = src._arrays
        np2 = (a['px'] ** 2 + a['py'] ** 2 + a['pz'] ** 2) / (a['mass'] * sc.c) ** 2
        ekin = np2 / (np.sqrt(1 + np2) + 1) * a['mass'] * sc.c ** 2 / sc.elementary_charge / 1e6
        h, ex, ey = cf.histogram2d(a['x'], a['y'], weights=ekin, range=((-3, 3), (-3, 3)), bins=(64, 64), shape=3)
        dV = (ex[1] - ex[0]) * (ey[1] - ey[0])
        assert np.allclose(f.matrix, h / dV, rtol=1e-11, atol=0)
        assert f.name.startswith('Ekin_MeV')
        pd = MultiSpecies(self._species(device='cuda'))
        fd = pd.createField('x', 'y', **kw)
        # row f4: the finished grid stays in HBM until the host matrix is first read
        assert f.device_matrix is None
        assert fd._matrix is None and fd.device_matrix.is_cuda and fd.shape == (64, 64) and fd.dimensions == 2
        assert np.allclose(fd.device_matrix.cpu().numpy(), f.matrix, rtol=1e-10, atol=1e-300)
        assert np.allclose(fd.matrix, f.matrix, rtol=1e-10, atol=1e-300)
        assert fd._matrix is not None and np.asarray(fd) is fd.matrix
        h2, _ = pd._createHistgram('x', 'y', **kw)          # the (h, edges) contract itself stays numpy
        assert isinstance(h2, np.ndarray) and np.allclose(h2, fd.matrix, rtol=1e-12, atol=1e-300)

    @pytest.mark.parametrize("kind", ['gamma_m1', 'gamma', 'Ekin', 'Ekin_MeV', 'Ekin_keV', 'p'])
    def test_fused_kinetic_weights_match_the_expression_chain(self, kind):
        """ptg_weights_f64 (one pass) against the scalar definitions of scalarproperties.py:183-205
        evaluated with numpy; tolerance 1e-14 relative (same operations, same order)."""
        import scipy.constants as sc
        import torch
        from postpic_b200 import device
        rng = np.random.default_rng(5)
        n = 200001
        me = 9.109383e-31
        px, py, pz = (rng.normal(0, 1, n) * me * sc.c * 3 for _ in range(3))
        wt = rng.uniform(0.5, 2, n)
        np2 = (px ** 2 + py ** 2 + pz ** 2) / (me * sc.c) ** 2
        gm1 = np2 / (np.sqrt(1 + np2) + 1)
        expect = {'gamma_m1': gm1, 'gamma': gm1 + 1, 'Ekin': gm1 * me * sc.c ** 2,
                  'Ekin_MeV': gm1 * me * sc.c ** 2 / sc.elementary_charge / 1e6,
                  'Ekin_keV': gm1 * me * sc.c ** 2 / sc.elementary_charge / 1e3,
                  'p': np.sqrt(px ** 2 + py ** 2 + pz ** 2)}[kind]
        t = [torch.from_numpy(a).cuda() for a in (px, py, pz, wt)]
        got = device.kinetic_weights(kind, t[0], t[1], t[2], t[3], mass=me, c=sc.c,
                                     elementary_charge=sc.elementary_charge).cpu().numpy()
        assert np.allclose(got, wt * expect, rtol=1e-14, atol=0)
        got = device.kinetic_weights(kind, t[0][::2], t[1][::2], t[2][::2], None, mass=me, c=sc.c,
                                     elementary_charge=sc.elementary_charge).cpu().numpy()
        assert np.allclose(got, expect[::2], rtol=1e-14, atol=0)

    @pytest.mark.parametrize("method", [None, 'rows', 'global', 'tiled', 'runs', 'runs_wide', 'sorted'])
    def test_fused_deposit_equals_two_step(self, method):
        """ptg_deposit_fused_f64 (kinetic weight evaluated inside the deposit kernels) against
        ptg_weights_f64 followed by ptg_deposit_f64, for every method, on random and cell-ordered input."""
        import scipy.constants as sc
        import torch
        from postpic_b200 import device
        n = 3_000_000
        g = torch.Generator(device='cuda').manual_seed(8)
        x = torch.rand(n, dtype=torch.float64, device='cuda', generator=g) * 1.02 - 0.01
        y = torch.rand(n, dtype=torch.float64, device='cuda', generator=g) * 1.02 - 0.01
        me = 9.109383e-31
        px, py, pz = (torch.randn(n, dtype=torch.float64, device='cuda', generator=g) * me * sc.c for _ in range(3))
        wt = torch.rand(n, dtype=torch.float64, device='cuda', generator=g) + 0.5
        order = torch.argsort((x.clamp(0, 1) * 200).floor() * 200 + (y.clamp(0, 1) * 200).floor())
        for idx in (None, order):
            sel = (lambda t: t) if idx is None else (lambda t: t[idx].contiguous())
            for shape in (1, 2) if method != 'runs_wide' else (1,):
                for w_arr in (sel(wt), None):
                    fused = dict(kind='Ekin_MeV', px=sel(px), py=sel(py), pz=sel(pz), mass=me, c=sc.c,
                                 elementary_charge=sc.elementary_charge)
                    kw = dict(range=[(0, 1), (0, 1)], bins=(200, 200), shape=shape)
                    got, _ = device.deposit([sel(x), sel(y)], w_arr, method=method, fused=fused, **kw)
                    w2 = device.kinetic_weights('Ekin_MeV', fused['px'], fused['py'], fused['pz'], w_arr, mass=me,
                                                c=sc.c, elementary_charge=sc.elementary_charge)
                    ref, _ = device.deposit([sel(x), sel(y)], w2, method='global', **kw)
                    assert float((got - ref).abs().max()) <= 1e-12 * float(ref.abs().max())

    def test_multi_species_on_device_deposits_without_concatenation(self):
        from postpic_b200 import MultiSpecies, ArraySource
        import torch
        a, b = self._species(30000), self._species(20000)
        b._arrays['x'] = b._arrays['x'] + 0.5
        host = MultiSpecies(a, b)

        def dev(src):
            return ArraySource(src.species, {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v)
                                             for k, v in src._arrays.items()}, extent=src.extent)
        device_ms = MultiSpecies(dev(a), dev(b))
        for kw in (dict(bins=[64, 48], rangex=(-3, 3), rangey=(-3, 3), shape=2),
                   dict(bins=[64, 48], shape=3, weights='Ekin_MeV'),                 # ranges from the data
                   dict(bins=[32, 32], simextent=True, shape=1, weights='gamma')):
            fh = host.createField('x', 'y', **kw)
            fd = device_ms.createField('x', 'y', **kw)
            assert all(np.array_equal(p.grid_node, q.grid_node) for p, q in zip(fh.axes, fd.axes))
            assert np.allclose(fd.matrix, fh.matrix, rtol=1e-10, atol=1e-12 * np.abs(fh.matrix).max())
            assert fd.matrix.base is not None and fd.infostring == '50000 npart in 2 species'

    def test_filter_is_predicated_deposition_on_device(self):
        """MultiSpecies.filter: host species are compressed (reference behaviour), device species carry
        a mask into the kernels; both give the same field, also with fused energy weights."""
        from postpic_b200 import MultiSpecies, ArraySource
        import torch
        a = self._species(40000)
        host = MultiSpecies(a).filter('gamma > 1.2').filter('x > -1')

        def dev(src):
            return ArraySource(src.species, {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v)
                                             for k, v in src._arrays.items()}, extent=src.extent)
        device_ms = MultiSpecies(dev(a), dev(a)).filter('gamma > 1.2').filter('x > -1')
        assert 0 < host.npart < 40000 and device_ms.npart == 2 * host.npart
        for kw in (dict(bins=[48, 40], rangex=(-3, 3), rangey=(-3, 3), shape=2),
                   dict(bins=[48, 40], shape=1, weights='Ekin_MeV'),                # ranges from the selected particles
                   dict(bins=[30], shape=3, weights='gamma')):
            axes = ('x', 'y')[:len(kw['bins'])]
            fh = host.createField(*axes, **kw)
            fd = device_ms.createField(*axes, **kw)
            assert all(np.array_equal(p.grid_node, q.grid_node) for p, q in zip(fh.axes, fd.axes))
            assert np.allclose(fd.matrix, 2 * fh.matrix, rtol=1e-10, atol=1e-12 * np.abs(fh.matrix).max())
        none = MultiSpecies(dev(a)).filter('gamma > 1e9')
        f = none.createField('x', 'y', bins=[4, 4], rangex=(-1, 1), rangey=(-1, 1))
        assert f.matrix.shape == (4, 4) and not f.matrix.any()

    def test_dummy_species_host_and_device_agree(self):
        """test/test_particles.py:37-49 on the reference's own fixture: createField of p + p is twice that
        of p, for the dummy species on the host and uploaded to the device."""
        from postpic_b200 import MultiSpecies, ArraySource
        ph = MultiSpecies(ArraySource.dummy(10000, dimensions=2))
        pd = MultiSpecies(ArraySource.dummy(10000, dimensions=2, device='cuda:0'))
        for p in (ph, pd):
            f1 = p.createField('x', 'y', bins=[50, 40], simextent=True)
            f2 = (p + p).createField('x', 'y', bins=[50, 40], simextent=True)
            assert np.allclose(2 * f1.matrix, f2.matrix) and f1.extent == [-2.0, 10.0, -5.0, 5.0]
        fh = ph.createField('x', 'y', bins=[50, 40], simextent=True, weights='Ekin_MeV', shape=3)
        fd = pd.createField('x', 'y', bins=[50, 40], simextent=True, weights='Ekin_MeV', shape=3)
        assert np.allclose(fd.matrix, fh.matrix, rtol=1e-10, atol=1e-12 * np.abs(fh.matrix).max())

    def test_empty_and_aliases(self):
        from postpic_b200 import MultiSpecies, ArraySource
        empty = MultiSpecies(ArraySource('none', dict(x=np.empty(0), y=np.empty(0), weight=np.empty(0))))
        h, edges = empty._createHistgram('x', 'y', bins=[5, 6], rangex=(0, 1))
        assert h.shape == (5, 6) and not h.any() and np.array_equal(edges[1], np.linspace(0, 1, 7))
        p = MultiSpecies(self._species())
        h1, _ = p.createHistgram1d('x', bins=[32], rangex=(-3, 3))
        h2, _ = p.createHistgram2d('x', 'y', bins=[8, 8], rangex=(-3, 3), rangey=(-3, 3))
        h3, _ = p.createHistgram3d('x', 'y', 'z', bins=[4, 4, 4], rangex=(-3, 3), rangey=(-3, 3), rangez=(-3, 3))
        assert h1.shape == (32,) and h2.shape == (8, 8) and h3.shape == (4, 4, 4)
        with pytest.raises(TypeError):
            p.createField('x', 'y', 'z', 'px')
