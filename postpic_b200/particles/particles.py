"""The histogram dispatch of `MultiSpecies` -- and only that.

Re-states the contract of
    MultiSpecies._createHistgram   /root/reference/postpic/particles/particles.py:1015-1103
    MultiSpecies.createField       /root/reference/postpic/particles/particles.py:1105-1167
(plus the historic `createHistgram1d/2d/3d` names, which that checkout merged into
`_createHistgram`) on top of the CUDA `histogramdd`.  Everything else `MultiSpecies`
does in postpic (species bookkeeping, compress/filter, particle tracking, file
readers, the plotting `Field` with its 2000 lines of methods) is out of scope; the
small stand-ins here (`ArraySource`, `Field`, the scalar table) exist so the dispatch
can be exercised and tested without postpic installed.  When postpic is installed, a
maintainer swaps one import instead.

The keyword handling and the empty-histogram branch of `_createHistgram` and the tail of
`createField` follow the reference statement by statement (particles.py:1043-1095, :1150-1167):
that is API-contract glue a drop-in must reproduce (same keywords, defaults, exceptions and
names), not the hot path -- everything from the axis / weight evaluation down is this
package's own.

Per-particle scalars of device-resident species are evaluated INSIDE the deposit kernels when
they are names of postpic's default scalar context (enumerated kinds, `device.SCALAR_KINDS`);
arbitrary expression strings go through a small whitelisting evaluator (numpy for host arrays,
torch ops for CUDA tensors).
"""
import ast
import warnings

import numpy as np

from ._routines import histogramdd
from .. import device as _device

__all__ = ['MultiSpecies', 'Field', 'ArraySource']

try:                                     # the reference resolves unknown names in scipy.constants
    import scipy.constants as _sc       # (particles.py:303,:322)
except ImportError:                      # pragma: no cover
    _sc = None

# fallback table when scipy is absent (CODATA values of the names the default scalars use)
_CONSTANTS = {
    'c': 299792458.0,
    'elementary_charge': 1.602176634e-19,
    'atomic_mass': 1.66053906660e-27,
    'electron_mass': 9.1093837015e-31,
    'pi': np.pi,
}

# name -> (expression in terms of other scalars, unit).  The definitions are those of
# postpic's default scalar context (scalarproperties.py:159-222); written as data.
_SCALARS = {
    'w': ('weight', 'npartpermacro'), 'm': ('mass', 'kg'), 'q': ('charge', 'C'),
    'mass_u': ('mass / atomic_mass', 'amu'), 'm_u': ('mass / atomic_mass', 'amu'),
    'charge_e': ('charge / elementary_charge', ''), 'q_e': ('charge / elementary_charge', ''),
    'x_um': ('x * 1e6', r'$\mu$ m'), 'y_um': ('y * 1e6', r'$\mu$ m'), 'z_um': ('z * 1e6', r'$\mu$ m'),
    'p': ('sqrt(px**2 + py**2 + pz**2)', 'kg*m/s'),
    '_np2': ('(px**2 + py**2 + pz**2)/(mass * c)**2', ''),
    'gamma_m1': ('_np2 / (sqrt(1 + _np2) + 1)', ''),
    'gamma': ('_np2 / (sqrt(1 + _np2) + 1) + 1', ''),
    'gamma_m': ('gamma * mass', 'kg'),
    'beta': ('sqrt(gamma_m1**2 + 2 * gamma_m1) / (gamma_m1 + 1)', ''),
    'v': ('beta * c', 'm/s'),
    'vx': ('px / (gamma * mass)', 'm/s'), 'vy': ('py / (gamma * mass)', 'm/s'), 'vz': ('pz / (gamma * mass)', 'm/s'),
    'betax': ('vx / c', ''), 'betay': ('vy / c', ''), 'betaz': ('vz / c', ''),
    'Eruhe': ('mass * c**2', 'J'),
    'Ekin': ('gamma_m1 * mass * c**2', 'J'),
    'Ekin_MeV': ('Ekin / elementary_charge / 1e6', 'MeV'),
    'Ekin_MeV_amu': ('Ekin / elementary_charge / 1e6 / mass_u', 'MeV/u'),
    'Ekin_MeV_qm': ('Ekin / elementary_charge / 1e6 / mass_u * charge_e', 'MeV * q/m'),
    'Ekin_keV': ('Ekin / elementary_charge / 1e3', 'keV/u'),          # sic: the reference's unit string (:205)
    'Ekin_keV_amu': ('Ekin / elementary_charge / 1e3 / mass_u', 'keV/u'),
    'Ekin_keV_qm': ('Ekin / elementary_charge / 1e3 / mass_u * charge_e', 'keV * q/m'),
    'angle_xaxis': ('arctan2(sqrt(py**2 + pz**2), px)', 'rad'),
    'angle_yaxis': ('arctan2(sqrt(pz**2 + px**2), py)', 'rad'),
    'angle_zaxis': ('arctan2(sqrt(px**2 + py**2), pz)', 'rad'),
    't': ('time', 's'),
    'angle_xy': ('arctan2(py, px)', 'rad'), 'angle_yx': ('arctan2(px, py)', 'rad'),
    'angle_yz': ('arctan2(pz, py)', 'rad'), 'angle_zy': ('arctan2(py, pz)', 'rad'),
    'angle_zx': ('arctan2(px, pz)', 'rad'), 'angle_xz': ('arctan2(pz, px)', 'rad'),
    'r_xy': ('sqrt(x**2 + y**2)', 'm'), 'r_yz': ('sqrt(y**2 + z**2)', 'm'),
    'r_zx': ('sqrt(z**2 + x**2)', 'm'), 'r_xyz': ('sqrt(x**2 + y**2 + z**2)', 'm'),
}
_ATOMIC_UNITS = {'x': 'm', 'y': 'm', 'z': 'm', 'px': 'kg*m/s', 'py': 'kg*m/s', 'pz': 'kg*m/s',
                 'weight': 'npartpermacro', 'mass': 'kg', 'charge': 'C', 'id': '', 'time': 's'}


# short symbols whose ScalarProperty carries the long name (scalarproperties.py:161-170)
_SYMBOL_NAMES = {'w': 'weight', 'm': 'mass', 'q': 'charge', 'm_u': 'mass_u', 'q_e': 'charge_e', 't': 'time'}


def _unit_of(expr):
    if expr in _ATOMIC_UNITS:
        return _ATOMIC_UNITS[expr]
    if expr in _SCALARS:
        return _SCALARS[expr][1]
    return ''


def _findscalarattr(scalarf, attrib, default='unknown'):
    """`name` / `unit` of a scalar given as an object with that attribute or as the name of a known
    scalar; anything else -- e.g. an expression string -- is 'unknown' (particles.py:44-57)."""
    ret = getattr(scalarf, attrib, None) if not isinstance(scalarf, str) else None
    if isinstance(scalarf, str) and (scalarf in _SCALARS or scalarf in _ATOMIC_UNITS):
        ret = _SYMBOL_NAMES.get(scalarf, scalarf) if attrib == 'name' else _unit_of(scalarf)
    return default if ret is None else ret


# scalar kinds (device.SCALAR_KINDS) that read the species' mass / charge as a constant
_NEED_MASS = frozenset([1, 2, 3, 4, 5, 14, 17] + list(range(22, 37)))
_NEED_CHARGE = frozenset([15, 18, 34, 36])


class ArraySource(object):
    """One species worth of particles held as arrays (numpy, or torch CUDA tensors).

    Stand-in for a `_SingleSpecies` bound to a dumpreader (particles.py:60-340): it
    answers `atomic(name)` with a 1-D float64 array of length `len(self)`; scalars
    given as Python numbers (e.g. mass=9.1e-31, weight=1.0) are broadcast lazily.
    `extent` maps an axis name to (lo, hi) -- the reader's `simextent(axis)`.
    """

    def __init__(self, species, arrays, extent=None):
        self.species = species
        self._arrays = dict(arrays)
        self.extent = dict(extent or {})
        lens = [len(v) for v in self._arrays.values() if np.ndim(v) == 1 or _device.is_cuda_tensor(v)]
        if not lens:
            raise ValueError('at least one per-particle array is required')
        if len(set(lens)) != 1:
            raise ValueError('all per-particle arrays of a species must have equal length')
        self._n = lens[0]
        sample = [v for v in self._arrays.values() if _device.is_cuda_tensor(v)]
        self.device = sample[0].device if sample else None

    def __len__(self):
        return self._n

    @classmethod
    def dummy(cls, n, dimensions=2, seed=0, device=None):
        """The synthetic electron species of postpic's dummy reader (datareader/dummy.py:43-61, :133-176,
        :177-201), the reference's only particle fixture: np.random.seed(seed); x, y, z ~ N(0,1);
        px = roll(x,1)**2 * me*c, py = roll(x,2) * me*c, pz = roll(x,3) * me*c; weight 1; shuffled ids;
        simulation extent x (-2,10), y (-5,5), z (-4,4).  With `device` the arrays go straight into
        device buffers through the library's pinned staging ring."""
        me, c, qe = 9.109383e-31, 299792458.0, 1.602176565e-19       # helper.PhysicalConstants
        state = np.random.get_state()
        try:
            if seed is not None:
                np.random.seed(seed)
            n = int(n)
            arrs = {'x': np.random.normal(size=n)}
            if dimensions > 1:
                arrs['y'] = np.random.normal(size=n)
            if dimensions > 2:
                arrs['z'] = np.random.normal(size=n)
            x = arrs['x']
            arrs['px'] = np.roll(x, 1) ** 2 * (me * c)
            arrs['py'] = np.roll(x, 2) * (me * c)
            arrs['pz'] = np.roll(x, 3) * (me * c)
            ids = np.arange(n)
            np.random.shuffle(ids)
            arrs['id'] = ids.astype(np.float64)
        finally:
            np.random.set_state(state)
        if device is not None:
            idx = _device._torch().device(device).index
            arrs = {k: _device.upload(v, idx) for k, v in arrs.items()}
        arrs.update(weight=1.0, mass=me, charge=-qe, time=n * 1e-10)      # dummy.py:77-81: time = timestep * 1e-10
        extent = {'x': (-2.0, 10.0), 'y': (-5.0, 5.0), 'z': (-4.0, 4.0)}
        return cls('electron', arrs, extent={k: extent[k] for k in 'xyz'[:dimensions]})

    ATTRIBS = ('x', 'y', 'z', 'px', 'py', 'pz', 'weight', 'mass', 'charge', 'id')

    @classmethod
    def from_dumpreader(cls, reader, species, device=None, attribs=None, chunk=1 << 22):
        """One species of a postpic dump reader (anything with the particle half of `Dumpreader_ifc`,
        /root/reference/postpic/datareader/datareader.py:247-263: `getSpecies(species, attrib)` returning a
        per-particle sequence, a single float for a per-species constant, or raising KeyError; optionally
        `time()`, `simextent(axis)`, `simdimensions()`) as an ArraySource.

        Every attribute is pulled on its own and -- with `device` -- goes chunk by chunk (`chunk` particles at a
        time) through the library's pinned staging ring straight into its device array: whatever the reader hands
        out (float32, integer ids, strided views, lazily read HDF5 / SDF datasets that only materialise the slice
        asked for) is converted to float64 one chunk at a time, so no full-size host copy of the species is ever
        made, and nothing is concatenated (the reference: _SingleSpecies.__getitem__ particles.py:128-150 reads
        whole arrays, MultiSpecies.__call_sp :662-674 hstacks them).  Attributes the dump lacks are skipped."""
        arrs = {}
        n = None
        dev = None
        if device is not None:
            torch = _device._torch()
            dev = torch.device(device)
            if dev.index is None:
                dev = torch.device('cuda', torch.cuda.current_device())
        for name in (attribs or cls.ATTRIBS):
            try:
                val = reader.getSpecies(species, name)
            except KeyError:
                continue
            try:                                            # (never np.ndim(val): that would materialise a lazy dataset)
                m = len(val)
            except TypeError:                               # a plain number or a 0-d array: per-species constant
                arrs[name] = float(val)                     # (mass, charge, a uniform weight)
                continue
            if n is None:
                n = m
            elif m != n:
                raise ValueError('attribute {!r} of species {!r} has {} entries, expected {}'.format(name, species, m, n))
            if dev is None:
                arrs[name] = np.asarray(val, dtype=np.float64)
                continue
            out = _device._torch().empty(m, dtype=_device._torch().float64, device=dev)
            for i0 in range(0, m, int(chunk)):
                _device.upload_into(out, i0, np.asarray(val[i0:i0 + int(chunk)], dtype=np.float64))
            arrs[name] = out
        if n is None:
            raise KeyError('species {!r} has no per-particle attribute'.format(species))
        arrs.setdefault('weight', 1.0)
        if hasattr(reader, 'time'):
            try:
                arrs['time'] = float(reader.time())
            except Exception:
                pass
        extent = {}
        if hasattr(reader, 'simextent'):
            for ax in 'xyz':
                try:
                    e = reader.simextent(ax)
                    extent[ax] = (float(e[0]), float(e[1]))
                except Exception:
                    pass
        return cls(species, arrs, extent=extent)

    def atomic(self, name):
        if name not in self._arrays:
            raise KeyError(name)
        v = self._arrays[name]
        if _device.is_cuda_tensor(v):
            return v
        if np.ndim(v) == 0:
            if self.device is not None:
                torch = _device._torch()
                return torch.full((self._n,), float(v), dtype=torch.float64, device=self.device)
            return np.full(self._n, float(v), dtype=np.float64)
        return np.asarray(v, dtype=np.float64)


_ALLOWED_NODES = (ast.Expression, ast.BinOp, ast.UnaryOp, ast.Compare, ast.BoolOp, ast.Call, ast.Name, ast.Load,
                  ast.Constant, ast.Add, ast.Sub, ast.Mult, ast.Div, ast.Pow, ast.Mod, ast.FloorDiv, ast.USub, ast.UAdd,
                  ast.Invert, ast.Not, ast.And, ast.Or, ast.BitAnd, ast.BitOr, ast.BitXor, ast.Eq, ast.NotEq, ast.Lt,
                  ast.LtE, ast.Gt, ast.GtE)
_code_cache = {}


def _safe_eval(expr, ns, funcs):
    """Evaluate an arithmetic expression string over the names in `ns`.  Like numexpr (which the reference uses,
    particles.py:290-330) it accepts arithmetic, comparisons and calls of the known functions only: attribute
    access, subscripts, lambdas, comprehensions ... are rejected before anything is evaluated."""
    code = _code_cache.get(expr)
    if code is None:
        try:
            tree = ast.parse(str(expr), mode='eval')
        except SyntaxError as exc:
            raise ValueError('cannot parse scalar expression {!r}: {}'.format(expr, exc))
        for node in ast.walk(tree):
            if not isinstance(node, _ALLOWED_NODES):
                raise ValueError('scalar expression {!r}: {} is not allowed'.format(expr, type(node).__name__))
            if isinstance(node, ast.Call) and not (isinstance(node.func, ast.Name) and node.func.id in funcs
                                                   and not node.keywords):
                raise ValueError('scalar expression {!r}: only calls of {} are allowed'.format(expr, sorted(funcs)))
            if isinstance(node, ast.Constant) and not isinstance(node.value, (int, float, bool)):
                raise ValueError('scalar expression {!r}: only numeric constants are allowed'.format(expr))
        code = compile(tree, '<scalar>', 'eval')
        _code_cache[expr] = code
    return eval(code, {'__builtins__': {}}, ns)          # noqa: S307 (whitelisted AST)


class _Namespace(dict):
    """Name resolution for `eval`: atomic arrays, derived scalars (recursively), constants."""

    def __init__(self, source, funcs):
        dict.__init__(self)
        self._source = source
        self._funcs = funcs
        self.update(funcs)

    def __missing__(self, key):
        if key in _SCALARS:
            val = _safe_eval(_SCALARS[key][0], self, self._funcs)
        else:
            try:
                val = self._source.atomic(key)
            except KeyError:
                if _sc is not None and isinstance(getattr(_sc, key, None), float):
                    val = getattr(_sc, key)
                elif key in _CONSTANTS:
                    val = _CONSTANTS[key]
                else:
                    raise
        self[key] = val
        return val


def _numpy_funcs():
    return {k: getattr(np, k) for k in ('sqrt', 'arctan2', 'sin', 'cos', 'tan', 'exp', 'log', 'abs', 'where')}


def _torch_funcs():
    torch = _device._torch()
    return {'sqrt': torch.sqrt, 'arctan2': torch.atan2, 'sin': torch.sin, 'cos': torch.cos, 'tan': torch.tan,
            'exp': torch.exp, 'log': torch.log, 'abs': torch.abs, 'where': torch.where}


class _DeviceGrid(object):
    """A finished (scaled) ghost-padded grid that still lives in HBM, plus the slice selecting the
    returned bins.  `host()` is the one device->host copy of the whole path."""

    def __init__(self, padded, interior):
        self.padded = padded
        self.interior = interior

    @property
    def tensor(self):
        return self.padded[self.interior]

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    def host(self):
        return self.padded.cpu().numpy()[self.interior]


class Field(object):
    """Minimal result container with the constructor contract of postpic's `Field`
    (datahandling.py:561-612): `Field(matrix, xedges=..., yedges=..., zedges=...)`
    keeps `np.asarray(matrix)` and one edge array per axis.

    SURVEY.md section 8 row f4: a field created from device-resident species keeps its grid in HBM
    (`device_matrix`, a torch view of the padded grid) and copies it to the host only when `matrix`
    (or `np.asarray(field)`) is first read, so fields that are only summed, scaled or all-reduced on
    the device never cross PCIe."""

    class _Axis(object):
        def __init__(self, grid_node):
            self.grid_node = np.asarray(grid_node)
            self.name = ''
            self.unit = ''

        @property
        def grid(self):
            return 0.5 * (self.grid_node[1:] + self.grid_node[:-1])

        @property
        def extent(self):
            return [self.grid_node[0], self.grid_node[-1]]

    def __init__(self, matrix, xedges=None, yedges=None, zedges=None, name='', unit=''):
        if isinstance(matrix, _DeviceGrid):
            self._device_grid, self._matrix = matrix, None
        else:
            self._device_grid, self._matrix = None, np.asarray(matrix)
        self.axes = [Field._Axis(e) for e in (xedges, yedges, zedges) if e is not None]
        if len(self.axes) != len(self.shape):
            raise ValueError('one edge array per matrix dimension is required')
        self.name = name
        self.unit = unit
        self.label = ''
        self.infostring = ''
        self.infos = []

    @property
    def matrix(self):
        if self._matrix is None:
            self._matrix = self._device_grid.host()
        return self._matrix

    @matrix.setter
    def matrix(self, value):
        self._device_grid, self._matrix = None, np.asarray(value)

    @property
    def device_matrix(self):
        """The histogram as a torch CUDA view (None for fields created from host arrays)."""
        return None if self._device_grid is None else self._device_grid.tensor

    @property
    def shape(self):
        return self._device_grid.shape if self._matrix is None else self._matrix.shape

    @property
    def dimensions(self):
        return len(self.shape)

    @property
    def extent(self):
        return [v for ax in self.axes for v in ax.extent]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.matrix, dtype=dtype)


class MultiSpecies(object):
    """A collection of particle species exposing postpic's histogram entry points."""

    def __init__(self, *sources):
        self._sources = list(sources)
        self._filters = []          # scalar conditions set by filter(); evaluated per species on demand

    def filter(self, condition):
        """A new MultiSpecies restricted to the particles for which `condition` (an expression of the
        per-particle scalars, e.g. 'gamma > 2') is true (particles.py:561-582).  Host species are
        compressed like in the reference; device-resident species keep their arrays and carry the
        condition as a selection mask that the deposit kernels apply (predicated deposition)."""
        out = MultiSpecies(*self._sources)
        out._filters = self._filters + [condition]
        return out

    def _mask_of(self, source):
        """Boolean selection of one species under all filters, or None."""
        if not self._filters:
            return None
        single = MultiSpecies(source)
        mask = None
        for cond in self._filters:
            m = single(cond)
            mask = m if mask is None else (mask & m)
        return mask

    # -- the small part of the container API the dispatch below relies on ----------
    @property
    def npart(self):
        if not self._filters:
            return sum(len(s) for s in self._sources)
        return int(sum(int(self._mask_of(s).sum()) for s in self._sources))

    @property
    def nspecies(self):
        return len(self._sources)

    @property
    def species(self):
        names = sorted(set(s.species for s in self._sources))
        return ' '.join(names)

    def __len__(self):
        return self.npart

    def __add__(self, other):
        if self._filters or other._filters:
            raise NotImplementedError('adding filtered collections is not supported by this shim')
        return MultiSpecies(*(self._sources + other._sources))

    def getcompresslog(self):
        return {'all': []}

    def simextent(self, axis):
        """Combined extent of all species for `axis`, or None (particles.py:396-404)."""
        ext = [s.extent[axis] for s in self._sources if axis in s.extent]
        if not ext or len(ext) != len(self._sources):
            return None
        return np.asarray([min(e[0] for e in ext), max(e[1] for e in ext)])

    def simgridpoints(self, axis):
        """Always None: the reference's method computes a value but never returns it
        (particles.py:407-416), so `simgrid=True` only implies `simextent=True`."""
        return None

    def __call__(self, expr):
        """Evaluate a scalar expression per particle and concatenate over species
        (particles.py:626-674)."""
        if callable(expr):
            return expr(self)
        expr = getattr(expr, 'symbol', expr)
        parts = []
        for s in self._sources:
            funcs = _torch_funcs() if s.device is not None else _numpy_funcs()
            ns = _Namespace(s, funcs)
            try:
                val = _safe_eval(str(expr), ns, funcs)
            except NameError as exc:        # unknown scalar: the reference raises KeyError('"name" not found!')
                raise KeyError('"{}" not found!'.format(getattr(exc, 'name', None) or str(exc)))
            if np.ndim(val) == 0 and not _device.is_cuda_tensor(val):
                val = s.atomic('weight') * 0 + float(val) if 'weight' in s._arrays else np.full(len(s), float(val))
            if self._filters and not getattr(self, '_raw', False):
                val = val[self._mask_of(s)]          # compressed view of the species (reference behaviour)
            parts.append(val)
        if not parts:
            return np.empty(0, dtype=np.float64)
        if len(parts) == 1:
            return parts[0]
        if _device.is_cuda_tensor(parts[0]):
            return _device._torch().cat(parts)
        return np.hstack(parts)

    def _fusable(self, weights):
        """The extra weight is a kinetic scalar that the CUDA kernels can evaluate themselves."""
        key = getattr(weights, 'symbol', weights)
        ok = (isinstance(key, str) and key in _device.WEIGHT_KINDS and bool(self._sources) and
              all(s.device is not None and all(_device.is_cuda_tensor(s._arrays.get(k)) for k in ('px', 'py', 'pz'))
                  and s._arrays.get('mass', None) is not None and np.ndim(s._arrays.get('mass')) == 0
                  for s in self._sources))
        return key if ok else None

    def _weights(self, weights):
        """`weight * (weights)` per particle (particles.py:1097).  Device-resident species whose extra
        weight is one of the kinetic scalars (gamma, gamma_m1, Ekin, Ekin_MeV, Ekin_keV, p) are evaluated
        by one fused CUDA pass over px, py, pz, weight (ptg_weights_f64) instead of the chain of
        expression temporaries; everything else goes through the generic evaluator."""
        key = self._fusable(weights)
        if key is None:
            return self('weight * ({})'.format(weights))
        parts = []
        for s in self._sources:
            wt = s._arrays.get('weight', 1.0)
            wt_t = wt if _device.is_cuda_tensor(wt) else None
            out = _device.kinetic_weights(key, s._arrays['px'], s._arrays['py'], s._arrays['pz'], wt_t,
                                          mass=float(s._arrays['mass']),
                                          c=_sc.c if _sc is not None else _CONSTANTS['c'],
                                          elementary_charge=(_sc.elementary_charge if _sc is not None
                                                             else _CONSTANTS['elementary_charge']))
            if wt_t is None and float(wt) != 1.0:
                out *= float(wt)
            parts.append(out)
        return parts[0] if len(parts) == 1 else _device._torch().cat(parts)

    # -- the dispatch ------------------------------------------------------------------
    def _createHistgram(self, *sps, **kwargs):
        """(h, edges) of a 1-, 2- or 3-D deposit of this particle collection, divided by
        the cell volume.  Keywords: simextent, simgrid, rangex/rangey/rangez, weights
        (default '1'), force, bins, shape (particles.py:1015-1103)."""
        if 'optargsh' in kwargs:
            warnings.warn('keyword "optargsh" is deprecated. Use "bins" and "shape" '
                          'arguments directly on "createField".', category=DeprecationWarning)
            kwargs.update(kwargs.pop('optargsh'))
        simextent = kwargs.pop('simextent', False)
        simgrid = kwargs.pop('simgrid', False)
        ranges = [kwargs.pop('rangex', None), kwargs.pop('rangey', None), kwargs.pop('rangez', None)]
        weights = kwargs.pop('weights', '1')
        force = kwargs.pop('force', False)
        bins = kwargs.pop('bins', None)
        shape = kwargs.pop('shape', None)
        keep_on_device = kwargs.pop('_keep_on_device', False)      # createField only (row f4)
        if len(kwargs) > 0:
            raise TypeError("got an unexpected keyword argument {}'".format(kwargs))
        if len(sps) > 3:
            raise TypeError('Only 1D, 2D or 3D Histograms can be created.')

        if simgrid:
            simextent = True
        on_device = bool(self._sources) and all(s.device is not None for s in self._sources)
        on_host = bool(self._sources) and all(s.device is None for s in self._sources)
        several = sum(1 for s in self._sources if len(s) > 0) > 1
        if self._sources and not (on_device or on_host):
            raise TypeError('a MultiSpecies must hold either host (numpy) species or CUDA-resident species, not both')
        try:
            if (on_device or (on_host and several)) and self.npart > 0:
                data = None                       # evaluated per species below, nothing is concatenated
                for src in self._sources:         # (unknown scalars must still raise here -- or give the
                    if len(src) == 0:             #  empty histogram under force=True -- for EVERY species)
                        continue
                    probe = MultiSpecies(ArraySource(src.species, {k: (v[:1] if (np.ndim(v) == 1 or _device.is_cuda_tensor(v)) else v)
                                                                   for k, v in src._arrays.items()}))
                    for sp in sps:
                        probe(sp)
            else:
                data = [self(sp) for sp in sps]
        except KeyError:
            if not force:
                raise
            data = [[]]
        if simextent:
            for i, sp in enumerate(sps):
                tmp = self.simextent(getattr(sp, 'symbol', sp))
                ranges[i] = tmp if tmp is not None else ranges[i]
        if simgrid:
            for i, sp in enumerate(sps):
                tmp = self.simgridpoints(getattr(sp, 'symbol', sp))
                if tmp is not None:
                    bins[i] = tmp
        if data is not None and len(data[0]) == 0:  # no particles: an all-zero histogram
            h = np.zeros(bins)

            def edges_of(r, nb):
                return np.linspace(r[0], r[1], nb + 1) if r is not None else np.linspace(0, 1, nb + 1)
            edges = [edges_of(r, bins[i]) for i, r in zip(range(h.ndim), ranges)]
            return h, edges

        if data is None and on_device:
            return self._createHistgram_per_species(sps, ranges, weights, bins, shape, keep_on_device)
        if data is None:
            return self._createHistgram_host_per_species(sps, ranges, weights, bins, shape)
        w = self._weights(weights)
        h, edges = histogramdd(data, weights=w, range=ranges, bins=bins, shape=shape)
        dV = np.prod([edge[1] - edge[0] for edge in edges])
        h /= dV
        return h, edges

    def _createHistgram_host_per_species(self, sps, ranges, weights, bins, shape):
        """Several host-resident species: one deposit per species, summed in the ghost-padded accumulator of
        the first -- deposition is linear in the particle set -- instead of `np.hstack`-ing every axis and
        weight array of all species first (particles.py:662-674).  Range ends that are not given are the
        minimum / maximum over all species, as they would be for the concatenated arrays."""
        from ._routines import _minmax
        nd = len(sps)
        per = []
        for src in self._sources:
            m = MultiSpecies(src)
            m._filters = list(self._filters)
            d = [np.asarray(m(sp), dtype='float64') for sp in sps]
            if len(d[0]) == 0:
                continue
            per.append((d, np.asarray(m._weights(weights), dtype='float64')))
        rng = []
        cache = {}
        for a in range(nd):
            r = ranges[a]
            lo, hi = (None, None) if r is None else (r[0], r[1])
            if lo is None or hi is None:
                mm = np.array([_minmax(d[a], cache) for d, _ in per])       # np.min / np.max propagate NaN
                lo = np.min(mm[:, 0]) if lo is None else lo
                hi = np.max(mm[:, 1]) if hi is None else hi
            rng.append((lo, hi))
        h = edges = None
        for d, w in per:
            hk, ek = histogramdd(d, weights=w, range=rng, bins=bins, shape=shape)
            if h is None:
                h, edges = hk, ek
            else:
                acc = h.base if h.base is not None else h
                acc += hk.base if hk.base is not None else hk
        dV = np.prod([edge[1] - edge[0] for edge in edges])
        h /= dV
        return h, edges

    def _scalar_plan(self, src, sps, weights):
        """(axis scalar names, weight expression name or None) when every axis and the extra weight of species
        `src` is a NAMED scalar of the default context whose atomic arrays are CUDA tensors -- then the deposit
        kernels evaluate them per particle themselves (device.deposit_scalars_into) and no N-sized temporary
        exists; None when anything needs the generic evaluator."""
        names = [getattr(sp, 'symbol', sp) for sp in sps]
        wkey = getattr(weights, 'symbol', weights)
        if wkey in ('1', 1, 1.0, None):
            wkey = None
        kinds = _device.SCALAR_KINDS
        if not all(isinstance(nm, str) and nm in kinds for nm in names + ([wkey] if wkey is not None else [])):
            return None
        arrs = src._arrays
        need = set()

        def number(key):
            v = arrs.get(key, None)
            return v is not None and not _device.is_cuda_tensor(v) and np.ndim(v) == 0
        for nm in names + ([wkey] if wkey is not None else []):
            k = kinds[nm]
            need.update(_device.scalar_arrays_needed(k))
            if k in _NEED_MASS and not number('mass'):
                return None                                   # per-particle or missing mass: generic evaluator
            if k in _NEED_CHARGE and not number('charge'):
                return None
            if k == 16 and not number('time'):
                return None
        if not all(_device.is_cuda_tensor(arrs.get(k)) for k in need):
            return None
        wt = arrs.get('weight', 1.0)
        if not (_device.is_cuda_tensor(wt) or np.ndim(wt) == 0):
            return None
        return names, wkey

    def _species_dict(self, src):
        c = _sc.c if _sc is not None else _CONSTANTS['c']
        qe = _sc.elementary_charge if _sc is not None else _CONSTANTS['elementary_charge']
        amu = _sc.atomic_mass if _sc is not None else _CONSTANTS['atomic_mass']
        d = {k: v for k, v in src._arrays.items() if k in ('x', 'y', 'z', 'px', 'py', 'pz', 'weight', 'mass', 'charge', 'time')}
        d.setdefault('weight', 1.0)
        d.update(c=c, elementary_charge=qe, atomic_mass=amu)
        return d

    def _createHistgram_per_species(self, sps, ranges, weights, bins, shape, keep_on_device=False):
        """Several device-resident species: deposition is linear in the particle set, so every
        species is deposited into the SAME device grid, one after the other, instead of first
        concatenating all per-particle arrays (the np.hstack of particles.py:662-674).  Named scalars are
        evaluated inside the deposit kernel (no N-sized temporaries); the 1/dV scaling also happens on the
        device; only the finished grid crosses PCIe."""
        from .._setup import as_bins
        from ._routines import _DEFAULT_BINS
        torch = _device._torch()
        nd = len(sps)
        dev = self._sources[0].device
        c = _sc.c if _sc is not None else _CONSTANTS['c']
        qe = _sc.elementary_charge if _sc is not None else _CONSTANTS['elementary_charge']
        per = []            # one job per species with at least one selected particle
        for src in self._sources:
            if len(src) == 0:
                continue
            m = MultiSpecies(src)
            mk = self._mask_of(src)
            if mk is not None and not bool(mk.any()):
                continue                                   # nothing of this species passes the filter
            plan_s = self._scalar_plan(src, sps, weights)
            if plan_s is not None:
                per.append(dict(kind='scalars', src=src, names=plan_s[0], wkey=plan_s[1], mask=mk))
                continue
            d = [m(sp) for sp in sps]
            d = [a if _device.is_cuda_tensor(a) else torch.as_tensor(np.asarray(a, dtype='float64'), device=dev)
                 for a in d]
            kind = m._fusable(weights)
            if kind is not None:
                # weight * <kinetic scalar> is evaluated inside the deposit kernels (ptg_deposit_fused_f64)
                wt = src._arrays.get('weight', 1.0)
                if _device.is_cuda_tensor(wt):
                    w = wt
                elif float(wt) == 1.0:
                    w = None
                else:
                    w = torch.full((len(src),), float(wt), dtype=torch.float64, device=dev)
                fused = dict(kind=kind, px=src._arrays['px'], py=src._arrays['py'], pz=src._arrays['pz'],
                             mass=float(src._arrays['mass']), c=c, elementary_charge=qe)
            else:
                w, fused = m._weights(weights), None
            per.append(dict(kind='arrays', d=d, w=w, fused=fused, mask=mk))
        if bins is None:
            bins = _DEFAULT_BINS[nd]
        elif np.isscalar(bins):
            bins = (bins,) * nd
        if not per:                                        # every species filtered away: an all-zero histogram
            h = np.zeros(tuple(as_bins(bins[0] if nd == 1 else bins, nd)))
            edges = tuple(np.linspace(r[0], r[1], nb + 1) if r is not None and r[0] is not None and r[1] is not None
                          else np.linspace(0, 1, nb + 1) for r, nb in zip(ranges[:nd], h.shape))
            return h, edges

        def axis_values(job, a):
            if job['kind'] == 'arrays':
                t = job['d'][a]
            else:                                          # a temporary, only to complete a missing range end
                t = _device.scalar(job['names'][a], self._species_dict(job['src']), len(job['src']), dev)
            return t if job['mask'] is None else t[job['mask']]
        rng = []
        for a in range(nd):
            r = ranges[a]
            lo, hi = (None, None) if r is None else (r[0], r[1])
            if lo is None or hi is None:
                mm = np.array([_device.minmax(axis_values(job, a)) for job in per])
                lo = np.min(mm[:, 0]) if lo is None else lo        # np.min / np.max propagate NaN like the
                hi = np.max(mm[:, 1]) if hi is None else hi        # reference's reduction over the hstack-ed data
            rng.append((lo, hi))
        plan = _device.Plan(rng, as_bins(bins[0] if nd == 1 else bins, nd), 2 if shape is None else shape)
        grid = torch.zeros(plan.dims, dtype=torch.float64, device=dev)
        for job in per:
            if job['kind'] == 'scalars':
                _device.deposit_scalars_into(grid, plan, self._species_dict(job['src']), len(job['src']), job['names'],
                                             weight_expr=job['wkey'], zero=False, mask=job['mask'])
                continue
            d, w = job['d'], job['w']
            _device.deposit_into(grid, plan, [t if t.dtype == torch.float64 else t.to(torch.float64) for t in d],
                                 w if (w is None or w.dtype == torch.float64) else w.to(torch.float64), zero=False,
                                 fused=job['fused'], mask=job['mask'])
        _device.divide_(grid, plan.cell_volume())          # h /= dV, an IEEE division like the reference's
        if keep_on_device:
            return _DeviceGrid(grid, plan.interior), tuple(plan.edges)
        h = grid.cpu().numpy()[plan.interior]
        return h, tuple(plan.edges)

    def createHistgram1d(self, spx, **kwargs):
        return self._createHistgram(spx, **kwargs)

    def createHistgram2d(self, spx, spy, **kwargs):
        return self._createHistgram(spx, spy, **kwargs)

    def createHistgram3d(self, spx, spy, spz, **kwargs):
        return self._createHistgram(spx, spy, spz, **kwargs)

    def createField(self, *sps, **kwargs):
        """n-d histogram wrapped in a `Field` with names/units filled in
        (particles.py:1105-1167).  Extra keywords: name (default 'distfn'), title."""
        name = kwargs.pop('name', 'distfn')
        title = kwargs.pop('title', None)
        h, edges = self._createHistgram(*sps, _keep_on_device=True, **kwargs)
        edgekw = {k: e for k, e in zip(['xedges', 'yedges', 'zedges'], edges)}
        ret = Field(h, **edgekw)
        if 'weights' in kwargs:
            name = _findscalarattr(kwargs['weights'], 'name')
        ret.name = str(name) + self.species
        ret.label = self.species
        ret.name = title if title else ret.name
        for i, sp in enumerate(sps):
            ret.axes[i].unit = _findscalarattr(sp, 'unit')
            ret.axes[i].name = _findscalarattr(sp, 'name')
        ret.infostring = '{:.0f} npart in {:.0f} species'.format(self.npart, self.nspecies)
        ret.infos = self.getcompresslog()['all']
        return ret
