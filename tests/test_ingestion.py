"""reader-shaped ingestion.  `ArraySource.from_dumpreader` takes anything with the particle
half of postpic's Dumpreader_ifc (/root/reference/postpic/datareader/datareader.py:247-263) and pulls the species'
attributes one by one, chunk-wise; the fake readers below mimic the reference's dummy reader
(/root/reference/postpic/datareader/dummy.py:43-61, :177-201) and the shapes real readers hand out (float32 positions,
integer ids, strided views, lazily sliced datasets, per-species constants)."""
import numpy as np
import pytest

ME, C0, QE = 9.109383e-31, 299792458.0, 1.602176565e-19


class FakeDummyReader(object):
    """Same data as the reference's Dummyreader(n, dimensions, seed)."""

    def __init__(self, n, dimensions=2, seed=0):
        state = np.random.get_state()
        np.random.seed(seed)
        self._d = {'x': np.random.normal(size=n)}
        if dimensions > 1:
            self._d['y'] = np.random.normal(size=n)
        if dimensions > 2:
            self._d['z'] = np.random.normal(size=n)
        x = self._d['x']
        self._d['px'] = np.roll(x, 1) ** 2 * (ME * C0)
        self._d['py'] = np.roll(x, 2) * (ME * C0)
        self._d['pz'] = np.roll(x, 3) * (ME * C0)
        self._d['weight'] = np.repeat(1, n)                       # an integer array, like dummy.py:58
        ids = np.arange(n)
        np.random.shuffle(ids)
        self._d['id'] = ids
        np.random.set_state(state)
        self._dims = dimensions
        self._n = n

    def listSpecies(self):
        return ['electron']

    def simdimensions(self):
        return self._dims

    def time(self):
        return self._n * 1e-10

    def simextent(self, axis):
        ext = {'x': (-2, 10), 'y': (-5, 5), 'z': (-4, 4)}
        if 'xyz'.index(axis) >= self._dims:
            raise KeyError(axis)
        return np.asarray(ext[axis], dtype=np.float64)

    def getSpecies(self, species, attrib):
        if species != 'electron':
            raise KeyError(species)
        if attrib == 'mass':
            return ME
        if attrib == 'charge':
            return -QE
        if attrib not in self._d:
            raise KeyError('Attrib "%s" of species "%s" not present' % (attrib, species))
        return self._d[attrib]


class LazyColumn(object):
    """A dataset that only materialises the slice asked for (h5py / sdf style) and remembers the largest one."""

    def __init__(self, arr):
        self._a = arr
        self.largest = 0

    def __len__(self):
        return len(self._a)

    def __getitem__(self, sl):
        out = self._a[sl]
        self.largest = max(self.largest, len(out) if np.ndim(out) else 1)
        return out


class AwkwardReader(FakeDummyReader):
    """float32 positions, strided momenta, lazily sliced weights."""

    def __init__(self, n):
        FakeDummyReader.__init__(self, n, 2, seed=5)
        self._d['x'] = self._d['x'].astype(np.float32)
        big = np.zeros(3 * n)
        big[::3] = self._d['px']
        self._d['px'] = big[::3]                                    # stride 3
        self._d['py'] = self._d['py'][::-1][::-1]
        self.lazy = LazyColumn(np.random.default_rng(1).random(n))
        self._d['weight'] = self.lazy


def test_from_dumpreader_equals_the_dummy_fixture_on_host():
    from postpic_b200 import ArraySource, MultiSpecies
    n = 5000
    a = ArraySource.from_dumpreader(FakeDummyReader(n, 2, seed=0), 'electron')
    b = ArraySource.dummy(n, dimensions=2, seed=0)
    for k in ('x', 'y', 'px', 'py', 'pz', 'id'):
        assert np.array_equal(a.atomic(k), b.atomic(k)), k
    assert np.array_equal(a.atomic('weight'), np.ones(n)) and a.atomic('weight').dtype == np.float64
    assert a._arrays['mass'] == ME and a._arrays['charge'] == -QE and a._arrays['time'] == n * 1e-10
    assert a.extent == {'x': (-2.0, 10.0), 'y': (-5.0, 5.0)}
    with pytest.raises(KeyError):
        a.atomic('z')                                               # a 2-D dump has no z (dummy.py:183-184)
    with pytest.raises(KeyError):
        ArraySource.from_dumpreader(FakeDummyReader(10), 'proton')
    assert np.array_equal(MultiSpecies(a)('gamma'), MultiSpecies(b)('gamma'))


@pytest.mark.gpu
def test_from_dumpreader_streams_chunks_into_device_arrays():
    import torch
    from postpic_b200 import ArraySource, MultiSpecies
    n = 300000
    rd = AwkwardReader(n)
    src = ArraySource.from_dumpreader(rd, 'electron', device='cuda', chunk=65536)
    assert rd.lazy.largest <= 65536                                 # never asked for more than one chunk
    host = ArraySource.from_dumpreader(AwkwardReader(n), 'electron')
    for k in ('x', 'y', 'px', 'py', 'pz', 'weight', 'id'):
        t = src._arrays[k]
        assert t.is_cuda and t.dtype == torch.float64 and t.numel() == n
        assert np.array_equal(t.cpu().numpy(), host.atomic(k)), k
    kw = dict(bins=[40, 30], shape=2, simextent=True, weights='Ekin_MeV')
    fd = MultiSpecies(src).createField('x', 'y', **kw)
    fh = MultiSpecies(host).createField('x', 'y', **kw)
    ref = np.asarray(fh.matrix)
    assert fd.device_matrix is not None
    assert np.max(np.abs(np.asarray(fd.matrix) - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_from_dumpreader_accepts_awkward_attribute_types_on_host():
    from postpic_b200 import ArraySource
    rd = AwkwardReader(4000)
    a = ArraySource.from_dumpreader(rd, 'electron')
    assert a.atomic('x').dtype == np.float64 and np.array_equal(a.atomic('x'), rd._d['x'].astype(np.float64))
    assert np.array_equal(a.atomic('px'), np.ascontiguousarray(rd._d['px']))
    assert np.array_equal(a.atomic('weight'), rd.lazy._a)
