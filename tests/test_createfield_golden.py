"""`MultiSpecies._createHistgram` / `createField` against golden vectors
produced by the REAL postpic run in the build container (tests/golden/make_createfield_golden.py:
its own createField, histogramdd, scalar expressions and compiled _particlestogrid on its own dummy reader).

  * CPU (`not gpu`): this package's host stack with the oracle in place of the CUDA module -- checks the host
    logic (scalar evaluation, weights, ranges, filters, species concatenation, 1/dV, names/units) bit-tightly;
  * GPU: the same calls through the CUDA library, host-resident and device-resident species.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_createfield_golden import CASES, NPART          # noqa: E402  (case table only; no postpic import)

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "createfield_golden.npz"))


def _collection(dims, cond, doubled, device=None):
    from postpic_b200 import MultiSpecies, ArraySource
    ms = MultiSpecies(ArraySource.dummy(NPART, dimensions=dims, device=device))
    if doubled:
        ms = ms + ms            # (the shim adds before filtering; no case does both)
    if cond is not None:
        ms = ms.filter(cond)
    return ms


def _check(key, f, ms, rtol):
    ref = GOLDEN[key + "_matrix"]
    assert f.matrix.shape == ref.shape
    assert np.max(np.abs(f.matrix - ref)) <= rtol * np.max(np.abs(ref))
    for a, ax in enumerate(f.axes):
        assert np.allclose(ax.grid_node, GOLDEN[key + "_edges%d" % a], rtol=1e-15, atol=0)
        assert ax.name == str(GOLDEN[key + "_axname%d" % a])
        assert ax.unit == str(GOLDEN[key + "_axunit%d" % a])
    assert f.name == str(GOLDEN[key + "_name"])
    assert f.infostring == str(GOLDEN[key + "_infostring"])
    assert len(ms) == int(GOLDEN[key + "_npart"])


@pytest.fixture
def cpu_stack(monkeypatch, oracle):
    from postpic_b200.particles import _routines
    import postpic_b200.particles.particles as P
    monkeypatch.setattr(_routines, "ptg", oracle)
    monkeypatch.setattr(P, "histogramdd", _routines.histogramdd)
    return _routines


@pytest.mark.parametrize("key,dims,sps,kw,cond,doubled", CASES, ids=[c[0] for c in CASES])
def test_createfield_matches_real_postpic_on_cpu_stack(cpu_stack, key, dims, sps, kw, cond, doubled):
    ms = _collection(dims, cond, doubled)
    f = ms.createField(*sps, **{k: (list(v) if isinstance(v, list) else v) for k, v in kw.items()})
    _check(key, f, ms, rtol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("resident", ["host", "device"])
@pytest.mark.parametrize("key,dims,sps,kw,cond,doubled", CASES, ids=[c[0] for c in CASES])
def test_createfield_matches_real_postpic_on_gpu(key, dims, sps, kw, cond, doubled, resident):
    from postpic_b200.particles import _particlestogrid
    _particlestogrid.method = None
    ms = _collection(dims, cond, doubled, device=("cuda:0" if resident == "device" else None))
    f = ms.createField(*sps, **{k: (list(v) if isinstance(v, list) else v) for k, v in kw.items()})
    _check(key, f, ms, rtol=1e-12)


# ---- row a5: the call forms of histogramdd against the real postpic's histogramdd ----
from make_createfield_golden import hdd_inputs          # noqa: E402

HDD_KEYS = sorted(hdd_inputs().keys())


def _check_hdd(key, h, edges, rtol, exact_counts):
    ref = GOLDEN["hdd_" + key + "_h"]
    assert h.shape == ref.shape and h.dtype == np.float64
    if exact_counts:
        assert np.array_equal(h, ref)
    else:
        assert np.max(np.abs(h - ref)) <= rtol * np.max(np.abs(ref))
    for a, e in enumerate(edges):
        assert np.array_equal(e, GOLDEN["hdd_" + key + "_edges%d" % a])


def _is_count(key):
    kw = hdd_inputs()[key][1]
    return kw.get('shape', 2) == 0 and 'weights' not in kw


@pytest.mark.parametrize("key", HDD_KEYS)
def test_histogramdd_matches_real_postpic_on_cpu_stack(cpu_stack, key):
    data, kw = hdd_inputs()[key]
    h, edges = cpu_stack.histogramdd(data, **kw)
    _check_hdd(key, h, edges, 1e-14, _is_count(key))


@pytest.mark.gpu
@pytest.mark.parametrize("key", HDD_KEYS)
def test_histogramdd_matches_real_postpic_on_gpu(key):
    from postpic_b200 import histogramdd
    from postpic_b200.particles import _particlestogrid
    _particlestogrid.method = None
    data, kw = hdd_inputs()[key]
    h, edges = histogramdd(data, **kw)
    _check_hdd(key, h, edges, 1e-12, _is_count(key))


# ---- the kinetic scalars of scalarproperties.py:183-205, evaluated by the reference's own expressions ----
from make_createfield_golden import SCALAR_KINDS, SCALAR_N          # noqa: E402


@pytest.mark.parametrize("kind", SCALAR_KINDS)
def test_scalar_expressions_match_real_postpic(kind):
    from postpic_b200 import MultiSpecies, ArraySource
    ms = MultiSpecies(ArraySource.dummy(SCALAR_N, dimensions=2))
    ref = GOLDEN["scalar_" + kind]
    got = np.asarray(ms(kind), dtype=np.float64)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= 4e-16 * np.max(np.abs(ref))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [k for k in SCALAR_KINDS if k not in ('beta', 'v')])
def test_fused_kinetic_weights_match_real_postpic(kind):
    """ptg_weights_f64 (the one-pass CUDA evaluation that createField fuses into the deposit) against the
    values the reference's scalar-property expressions give; 1e-14 relative (same operations, same order)."""
    import torch
    from postpic_b200 import ArraySource, device
    src = ArraySource.dummy(SCALAR_N, dimensions=2)
    a = src._arrays
    px, py, pz = (torch.from_numpy(np.ascontiguousarray(a[k])).cuda() for k in ('px', 'py', 'pz'))
    import scipy.constants as sc                     # the reference's scalar context takes c and e from scipy
    out = device.kinetic_weights(kind, px, py, pz, None, mass=float(a['mass']), c=sc.c,
                                 elementary_charge=sc.elementary_charge)
    ref = GOLDEN["scalar_" + kind]
    assert np.max(np.abs(out.cpu().numpy() - ref)) <= 1e-14 * np.max(np.abs(ref))


def test_whole_scalar_context_matches_real_postpic():
    """Every scalar of the reference's default context (scalarproperties.py:159-222): same name, same unit,
    same per-particle value on the dummy species (the ones the 2-D dummy reader cannot provide raise KeyError
    in both)."""
    from postpic_b200 import MultiSpecies, ArraySource
    from postpic_b200.particles.particles import _findscalarattr
    ms = MultiSpecies(ArraySource.dummy(SCALAR_N, dimensions=2))
    table = GOLDEN["ctx_table"]
    assert len(table) == 56
    for key, name, unit, available in table:
        assert _findscalarattr(str(key), 'name') == str(name), key
        assert _findscalarattr(str(key), 'unit') == str(unit), key
        if available == '1':
            ref = GOLDEN["ctx_" + str(key)]
            got = np.asarray(ms(str(key)), dtype=np.float64)
            scale = np.max(np.abs(ref))
            assert got.shape == ref.shape and np.max(np.abs(got - ref)) <= 1e-15 * scale + 0.0, key
        else:
            with pytest.raises(KeyError):
                ms(str(key))


def test_field_and_axis_naming_rules(cpu_stack):
    """Expression strings are 'unknown' (name and unit), aliases carry the long name (particles.py:44-57)."""
    from postpic_b200 import MultiSpecies, ArraySource
    ms = MultiSpecies(ArraySource.dummy(SCALAR_N, dimensions=2))
    f = ms.createField('x_um', 'q_e', bins=[4, 4], weights='gamma*2', rangex=(-3e6, 3e6), rangey=(-2, 0))
    assert [f.name, f.axes[0].name, f.axes[0].unit, f.axes[1].name, f.axes[1].unit] == [str(v) for v in GOLDEN["naming"]]
    f = ms.createField('px*2', bins=[4], rangex=(0, 1e-21))
    assert [f.name, f.axes[0].name, f.axes[0].unit] == [str(v) for v in GOLDEN["naming_expr_axis"]]
    with pytest.raises(KeyError, match='"zzz" not found!'):
        ms.createField('x', 'zzz', bins=[4, 4])


@pytest.mark.gpu
def test_every_enumerated_scalar_kind_matches_real_postpic():
    """ptg_scalar_f64 -- the device function the deposit kernels call for fused axis / weight expressions -- against
    the per-particle values of the reference's scalar context on its dummy species, every available scalar."""
    import torch
    from postpic_b200 import ArraySource, MultiSpecies, device
    src = ArraySource.dummy(SCALAR_N, dimensions=2, device='cuda')
    sp = MultiSpecies(src)._species_dict(src)
    checked = 0
    for key, name, unit, available in GOLDEN["ctx_table"]:
        key = str(key)
        if available != '1' or key not in device.SCALAR_KINDS:
            continue
        ref = GOLDEN["ctx_" + key]
        got = device.scalar(key, sp, len(src), torch.device('cuda', torch.cuda.current_device())).cpu().numpy()
        scale = np.max(np.abs(ref))
        tol = 4e-16 if key.startswith('angle') else 1e-15          # atan2: <= 2 ulp between libm and CUDA
        assert np.max(np.abs(got - ref)) <= tol * scale, key
        checked += 1
    assert checked >= 40


@pytest.mark.gpu
@pytest.mark.parametrize("axes,weights", [(('x_um', 'angle_xy'), 'gamma'), (('px', 'gamma'), None), (('r_xy',), 'Ekin_MeV'),
                                          (('x', 'y'), 'beta'), (('vx', 'vy'), 'w'), (('angle_xaxis', 'p'), 'Ekin_keV')])
def test_fused_axis_expressions_equal_eager_evaluation(axes, weights):
    """createField on device-resident species with named scalars on the axes: the kernels evaluate them per particle
    (no temporaries); the result must equal the deposit of the eagerly evaluated arrays, and the host (numpy) stack."""
    import torch
    from postpic_b200 import ArraySource, MultiSpecies, device
    n = 300000
    dsrc = ArraySource.dummy(n, dimensions=2, seed=3, device='cuda')
    hsrc = ArraySource.dummy(n, dimensions=2, seed=3)
    dms, hms = MultiSpecies(dsrc), MultiSpecies(hsrc)
    assert dms._scalar_plan(dsrc, axes, weights if weights else '1') is not None
    kw = dict(bins=[48, 40][:len(axes)], shape=2)
    for a, nm in zip('xyz', axes):
        v = np.asarray(hms(nm))
        kw['range' + a] = (float(v.min()) - 0.01 * abs(float(v.min())) - 1e-30, float(v.max()) * 1.01 + 1e-30)
    if weights:
        kw['weights'] = weights
    fd = dms.createField(*axes, **kw)
    fh = hms.createField(*axes, **kw)
    ref = np.asarray(fh.matrix)
    assert np.max(np.abs(np.asarray(fd.matrix) - ref)) <= 1e-11 * np.max(np.abs(ref))
    # with a filter (predicated deposition) and a range end taken from the data
    kw.pop('rangex')
    fd = dms.filter('gamma > 1.2').createField(*axes, **kw)
    fh = hms.filter('gamma > 1.2').createField(*axes, **kw)
    ref = np.asarray(fh.matrix)
    assert np.allclose(fd.axes[0].grid_node, fh.axes[0].grid_node, rtol=1e-14, atol=0)
    assert np.max(np.abs(np.asarray(fd.matrix) - ref)) <= 1e-11 * np.max(np.abs(ref))


@pytest.mark.gpu
def test_filtered_collection_with_an_empty_species_selection():
    """electrons + 'ions' under a filter no ion passes, range ends from the data: the reference histograms
    the concatenated compressed arrays without error; so must the per-species device path (and the host path)."""
    from postpic_b200 import ArraySource, MultiSpecies
    n = 20000
    for dev in (None, 'cuda'):
        e = ArraySource.dummy(n, dimensions=2, seed=1, device=dev)
        ion = ArraySource.dummy(n // 2, dimensions=2, seed=2, device=dev)
        for k in ('px', 'py', 'pz'):
            ion._arrays[k] = ion._arrays[k] * 1e-6              # slow: gamma ~ 1
        ion.species = 'ion'
        ms = MultiSpecies(e, ion).filter('gamma > 1.5')
        f = ms.createField('x', 'y', bins=[16, 12], shape=1)
        only = MultiSpecies(e).filter('gamma > 1.5').createField('x', 'y', bins=[16, 12], shape=1)
        assert np.allclose(np.asarray(f.matrix), np.asarray(only.matrix), rtol=1e-12, atol=0)
        assert np.array_equal(f.axes[0].grid_node, only.axes[0].grid_node)
