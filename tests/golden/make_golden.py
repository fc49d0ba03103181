#!/usr/bin/env python
"""Generate tests/golden/ptg_golden.npz by running the UNMODIFIED reference
(/root/reference/postpic/particles/_particlestogrid.pyx, built into oracle/_ref by
oracle/build_ref.py) on seeded inputs.  Run in the build container, where
/root/reference exists; the resulting small fixture is committed so that the GPU box
(which has no /root/reference) can check both the oracle and the CUDA path against it.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import build_ref  # noqa: E402
import cases  # noqa: E402

N = 4000


def main():
    if build_ref.build() is None:
        raise SystemExit("reference tree not available")
    from oracle._ref import _particlestogrid as ref
    out = {}
    for nd in (1, 2, 3):
        for shape in (0, 1, 2, 3):
            for weighted in (False, True):
                seed = 1000 + nd * 100 + shape * 10 + int(weighted)
                coords, w = cases.cloud(seed, N, nd, weighted)
                h, edges = cases.call(ref, coords, w, shape)
                key = "d%d_s%d_w%d" % (nd, shape, int(weighted))
                out[key + "_seed"] = np.int64(seed)
                out[key + "_padded"] = np.array(cases.padded(h))
                for a, e in enumerate(edges):
                    out[key + "_edges%d" % a] = e
    # special values on every axis kind: NaN, +-inf, huge, exact edges, just-outside values
    special = np.array([np.nan, np.inf, -np.inf, 1e300, -1e300, 3e9, -3e9, 0.0, -0.0, 1.0, 0.5,
                        -0.04, -0.06, -0.14, -0.16, 1.04, 1.06, 1.14, 1.16, 0.05, 0.95, 0.999999999,
                        np.nextafter(0.0, 1), np.nextafter(1.0, 0), np.nextafter(1.0, 2)])
    out["special_values"] = special
    good = np.full_like(special, 0.37)
    for shape in (0, 1, 2, 3):
        h, _ = ref.histogram(special, range=(0, 1), bins=10, shape=shape)
        out["special_d1_s%d" % shape] = np.array(cases.padded(h))
        # x special, y good (x is int-tested)
        h, _, _ = ref.histogram2d(special, good, range=((0, 1), (0, 1)), bins=(10, 7), shape=shape)
        out["special_d2x_s%d" % shape] = np.array(cases.padded(h))
        # y special only where the reference is defined: raw-double-tested axis must stay < 2^31
        ysafe = np.where(np.isfinite(special) & (np.abs(special) < 1e8), special, 0.37)
        ysafe = np.where(np.isnan(special) | (special < -1e8), special, ysafe)   # NaN, -inf, -huge are safe (fail u >= 0)
        out["special_ysafe"] = ysafe
        h, _, _ = ref.histogram2d(good, ysafe, range=((0, 1), (0, 1)), bins=(7, 10), shape=shape)
        out["special_d2y_s%d" % shape] = np.array(cases.padded(h))
        h, _, _, _ = ref.histogram3d(good, good, ysafe, range=((0, 1),) * 3, bins=(5, 6, 10), shape=shape)
        out["special_d3z_s%d" % shape] = np.array(cases.padded(h))
    # degenerate range lo == hi (pyx:53-57)
    d = np.array([2.0, 2.0, 2.0 + 4e-16, 1.5])
    for shape in (0, 1, 2, 3):
        h, e = ref.histogram(d, range=(2.0, 2.0), bins=4, shape=shape)
        out["degenerate_s%d" % shape] = np.array(cases.padded(h))
        out["degenerate_edges"] = e
    out["degenerate_data"] = d
    path = os.path.join(HERE, "ptg_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
