"""Pins the CPU oracle (oracle/ptg_oracle.c) -- the checker for the CUDA path -- against
(1) the committed golden vectors generated from the unmodified reference,
(2) the known-answer tables and boundary probes of SURVEY.md sections 4 and 8(a),
(3) the reference itself (oracle/_ref), bit for bit, when it is available,
(4) the reference's own property tests (test/test_particlestogrid.py) re-hosted.
No GPU needed."""
import os

import numpy as np
import pytest

import cases

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "ptg_golden.npz"))
ALL = [(nd, s, w) for nd in (1, 2, 3) for s in (0, 1, 2, 3) for w in (False, True)]


@pytest.mark.parametrize("nd,shape,weighted", ALL)
def test_oracle_matches_golden_bit_exact(oracle, nd, shape, weighted):
    key = "d%d_s%d_w%d" % (nd, shape, int(weighted))
    coords, w = cases.cloud(int(GOLDEN[key + "_seed"]), 4000, nd, weighted)
    h, edges = cases.call(oracle, coords, w, shape)
    assert np.array_equal(cases.padded(h), GOLDEN[key + "_padded"])
    for a, e in enumerate(edges):
        assert np.array_equal(e, GOLDEN[key + "_edges%d" % a])


@pytest.mark.parametrize("shape", [0, 1, 2, 3])
def test_oracle_special_values_golden(oracle, shape):
    sp = GOLDEN["special_values"]
    good = np.full_like(sp, 0.37)
    ysafe = GOLDEN["special_ysafe"]
    h, _ = oracle.histogram(sp, range=(0, 1), bins=10, shape=shape)
    assert np.array_equal(cases.padded(h), GOLDEN["special_d1_s%d" % shape])
    h, _, _ = oracle.histogram2d(sp, good, range=((0, 1), (0, 1)), bins=(10, 7), shape=shape)
    assert np.array_equal(cases.padded(h), GOLDEN["special_d2x_s%d" % shape])
    h, _, _ = oracle.histogram2d(good, ysafe, range=((0, 1), (0, 1)), bins=(7, 10), shape=shape)
    assert np.array_equal(cases.padded(h), GOLDEN["special_d2y_s%d" % shape])
    h, _, _, _ = oracle.histogram3d(good, good, ysafe, range=((0, 1),) * 3, bins=(5, 6, 10), shape=shape)
    assert np.array_equal(cases.padded(h), GOLDEN["special_d3z_s%d" % shape])


@pytest.mark.parametrize("shape", [0, 1, 2, 3])
def test_oracle_degenerate_range_golden(oracle, shape):
    h, e = oracle.histogram(GOLDEN["degenerate_data"], range=(2.0, 2.0), bins=4, shape=shape)
    assert np.array_equal(cases.padded(h), GOLDEN["degenerate_s%d" % shape])
    assert np.array_equal(e, GOLDEN["degenerate_edges"])


@pytest.mark.parametrize("shape", [0, 1, 2, 3])
def test_oracle_known_answers_1d(oracle, shape):
    h, e = oracle.histogram(cases.KNOWN_1D_POS, range=(0, 25), bins=25, shape=shape)
    expect = np.zeros(25)
    for k, v in cases.KNOWN_1D[shape].items():
        expect[k] = v
    np.testing.assert_allclose(h, expect, rtol=1e-15, atol=1e-17)
    assert np.array_equal(e, np.linspace(0, 25, 26))


def test_oracle_known_answer_2d(oracle):
    h, _, _ = oracle.histogram2d(np.array([2.25]), np.array([3.6]), weights=np.array([2.0]),
                                 range=((0, 5), (0, 6)), bins=(5, 6), shape=2)
    expect = np.zeros((5, 6))
    expect[1:4, 2:5] = cases.KNOWN_2D_BLOCK
    np.testing.assert_allclose(h, expect, rtol=1e-13, atol=1e-16)


@pytest.mark.parametrize("value", sorted(cases.BOUNDARY_PROBES))
def test_oracle_boundary_probes(oracle, value):
    for shape, (tot, inner) in enumerate(cases.BOUNDARY_PROBES[value]):
        h, _ = oracle.histogram(np.array([value]), range=(0, 1), bins=10, shape=shape)
        assert abs(cases.padded(h).sum() - tot) < 1e-12, (value, shape)
        assert abs(h.sum() - inner) < 5e-5, (value, shape)


def test_oracle_undefined_reference_cases_are_dropped(oracle):
    # raw-double-tested axes with values >= 2^31 / +inf crash the reference; defined here: dropped
    bad = np.array([1e300, np.inf, 3e9])
    good = np.full(3, 0.5)
    for shape in (1, 2, 3):
        h, _, _ = oracle.histogram2d(good, bad, range=((0, 1), (0, 1)), bins=(4, 4), shape=shape)
        assert cases.padded(h).sum() == 0.0
        h, _, _, _ = oracle.histogram3d(good, good, bad, range=((0, 1),) * 3, bins=(4, 4, 4), shape=shape)
        assert cases.padded(h).sum() == 0.0
    with pytest.raises(ValueError):
        oracle.histogram(good, range=(0, 1), bins=4, shape=5)
    with pytest.raises(ValueError):
        oracle.histogram(good, range=(0, 1), bins=4, weights=np.ones(2), shape=1)


@pytest.mark.parametrize("nd,shape,weighted", ALL)
def test_oracle_matches_reference_bit_exact(oracle, ref, nd, shape, weighted):
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    coords, w = cases.cloud(77 + nd + 10 * shape, 50000, nd, weighted, spill=0.2)
    ho, eo = cases.call(oracle, coords, w, shape)
    hr, er = cases.call(ref, coords, w, shape)
    assert np.array_equal(cases.padded(ho), cases.padded(hr))
    assert ho.shape == hr.shape and ho.strides == hr.strides
    for a, b in zip(eo, er):
        assert np.array_equal(a, b)


def test_oracle_randomised_vs_reference_bit_exact(oracle, ref):
    """300 random problems (reversed / tiny / huge ranges, 1..48 bins, NaN / -inf, values on the edges):
    the restatement must reproduce the executed reference bit for bit."""
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(777)
    for trial in range(300):
        nd = int(rng.integers(1, 4))
        shape = int(rng.integers(0, 4))
        n = int(rng.choice([0, 1, 2, 33, 1000, 4000]))
        bins = [int(rng.integers(1, 49)) for _ in range(nd)]
        ranges, coords = [], []
        for a in range(nd):
            lo = float(rng.normal(0, 10))
            span = float(10.0 ** rng.uniform(-6, 6))
            hi = lo + span
            if rng.integers(0, 4) == 0:
                lo, hi = hi, lo
            ranges.append((lo, hi))
            c = rng.uniform(min(lo, hi) - 0.2 * span, max(lo, hi) + 0.2 * span, n)
            if n > 8:
                c[0], c[1], c[4], c[5] = np.nan, -np.inf, lo, hi
                if a == 0 or (a == 2 and shape != 1):          # +inf / huge are defined on int-tested axes only
                    c[2], c[3] = np.inf, 1e300
            coords.append(c)
        w = rng.normal(1.0, 2.0, n) if rng.integers(0, 2) else None
        ho, eo = cases.call(oracle, coords, w, shape, ranges=ranges, bins=bins)
        hr, er = cases.call(ref, coords, w, shape, ranges=ranges, bins=bins)
        assert np.array_equal(cases.padded(ho), cases.padded(hr), equal_nan=True), (trial, nd, shape, bins, ranges)
        assert all(np.array_equal(a, b) for a, b in zip(eo, er))


def test_oracle_strided_inputs(oracle, ref):
    base = np.random.default_rng(5).uniform(0, 1, (3000, 3))
    cols = [base[:, i] for i in range(3)]
    dense = [np.ascontiguousarray(c) for c in cols]
    for shape in (0, 1, 2, 3):
        a, _ = cases.call(oracle, cols, None, shape, ranges=[(0, 1)] * 3, bins=[9, 8, 7])
        b, _ = cases.call(oracle, dense, None, shape, ranges=[(0, 1)] * 3, bins=[9, 8, 7])
        assert np.array_equal(cases.padded(a), cases.padded(b))


# ---- the reference's own tests (test/test_particlestogrid.py:26-261), re-hosted on the oracle ----
class TestReferenceProperties:
    @pytest.fixture(autouse=True)
    def _data(self, oracle):
        rng = np.random.default_rng(2024)
        self.cf = oracle
        self.data = rng.random(int(1e5))
        self.weights = rng.random(int(1e5))
        self.x3 = rng.random(1000)
        self.y3 = 2 * rng.random(1000)
        self.z3 = 3 * rng.random(1000)
        self.w3 = rng.random(1000)

    def test_histogram_equals_numpy(self):
        cfh, cfe = self.cf.histogram(self.data, bins=20, range=(0, 1), shape=0)
        nph, npe = np.histogram(self.data, bins=20, range=(0, 1))
        assert list(nph) == list(cfh) and list(npe) == list(cfe)

    def test_histogramw_equals_numpy(self):
        cfh, cfe = self.cf.histogram(self.data, bins=100, range=(0, 1), shape=0, weights=self.weights)
        nph, npe = np.histogram(self.data, bins=100, range=(0, 1), weights=self.weights)
        assert np.sum(np.abs(cfh - nph)) < 1e-7 and list(npe) == list(cfe)

    @pytest.mark.parametrize("shape", [1, 2, 3])
    @pytest.mark.parametrize("weighted", [False, True])
    def test_mass_conservation_1d(self, shape, weighted):
        w = self.weights if weighted else None
        cfh, cfe = self.cf.histogram(self.data, bins=100, range=(0.0, 1.0), shape=shape, weights=w)
        total = np.sum(w) if weighted else len(self.data)
        assert abs(np.sum(cfh.base) - total) < 1e-7
        assert list(np.histogram(self.data, bins=100, range=(0.0, 1.0))[1]) == list(cfe)

    @pytest.mark.parametrize("shape", [0, 1, 2, 3])
    @pytest.mark.parametrize("weighted", [False, True])
    def test_mass_conservation_2d(self, shape, weighted):
        w = self.w3 if weighted else None
        cfh, ex, ey = self.cf.histogram2d(self.x3, self.y3, weights=w, bins=(20, 30),
                                          range=((0, 1), (0, 2)), shape=shape)
        total = np.sum(w) if weighted else 1000
        base = cfh.base if cfh.base is not None else cfh
        assert abs(np.sum(base) - total) < 1e-7
        _, npx, npy = np.histogram2d(self.x3, self.y3, bins=(20, 30), range=((0, 1), (0, 2)))
        assert list(npx) == list(ex) and list(npy) == list(ey)

    @pytest.mark.parametrize("shape", [0, 1, 2, 3])
    @pytest.mark.parametrize("weighted", [False, True])
    def test_mass_conservation_3d(self, shape, weighted):
        w = self.w3 if weighted else None
        cfh, ex, ey, ez = self.cf.histogram3d(self.x3, self.y3, self.z3, weights=w, bins=(20, 30, 40),
                                              range=((0, 1), (0, 2), (0, 3)), shape=shape)
        total = np.sum(w) if weighted else 1000
        base = cfh.base if cfh.base is not None else cfh
        assert abs(np.sum(base) - total) < 1e-7
