"""Worker of tests/test_integration_swap.py (own process: importing the real postpic installs stand-ins for numexpr /
future / matplotlib in sys.modules).  Executes INTEGRATION.md section 1: the REAL
/root/reference/postpic/particles/_routines.py with its `ptg` module swapped for postpic_b200's drop-in.  Without a
GPU the drop-in's device step (`_deposit`) is replaced by the CPU oracle, so what is exercised is the signature
contract: positional order, keywords, return shapes, strides and edge arrays as the reference's caller uses them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, ROOT)


def main():
    from make_createfield_golden import import_reference_postpic, hdd_inputs, CASES, NPART
    postpic = import_reference_postpic()
    import postpic.particles._routines as R
    from postpic_b200.particles import _particlestogrid as dropin
    golden = np.load(os.path.join(HERE, "golden", "createfield_golden.npz"), allow_pickle=False)
    use_gpu = False
    try:
        import torch
        use_gpu = torch.cuda.is_available()
    except Exception:
        pass
    if not use_gpu:
        from oracle import ptg_oracle

        def cpu_deposit(coords, names, weights, ranges, bins, shape):
            if len(coords) == 1:
                h, e = ptg_oracle.histogram(coords[0], range=ranges, bins=bins, weights=weights, shape=shape)
                return h, [e]
            if len(coords) == 2:
                h, ex, ey = ptg_oracle.histogram2d(coords[0], coords[1], weights=weights, range=ranges, bins=bins, shape=shape)
                return h, [ex, ey]
            h, ex, ey, ez = ptg_oracle.histogram3d(coords[0], coords[1], coords[2], weights=weights, range=ranges,
                                                   bins=bins, shape=shape)
            return h, [ex, ey, ez]
        dropin._deposit = cpu_deposit
    assert R.ptg is not dropin
    R.ptg = dropin                                   # the one-line swap of INTEGRATION.md
    # (1) the reference's histogramdd, every call form, on top of the drop-in
    for key, (data, kw) in sorted(hdd_inputs().items()):
        h, edges = R.histogramdd(data, **kw)
        ref = golden["hdd_" + key + "_h"]
        assert h.shape == ref.shape and h.dtype == np.float64, key
        assert np.max(np.abs(h - ref)) <= 1e-12 * np.max(np.abs(ref)), key
        for a, e in enumerate(edges):
            assert np.array_equal(e, golden["hdd_" + key + "_edges%d" % a]), key
    # (2) the reference's MultiSpecies.createField on its dummy reader, on top of the drop-in
    postpic.chooseCode('dummy')
    for key, dims, sps, kw, cond, doubled in CASES:
        ms = postpic.MultiSpecies(postpic.readDump(NPART, dimensions=dims), 'electron')
        if cond:
            ms = ms.filter(cond)
        if doubled:
            ms = ms + ms
        f = ms.createField(*sps, **{k: (list(v) if isinstance(v, list) else v) for k, v in kw.items()})
        ref = golden[key + "_matrix"]
        got = np.asarray(f.matrix)
        assert got.shape == ref.shape, key
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref)), key
    print("INTEGRATION_SWAP_OK gpu=%s" % use_gpu)


if __name__ == "__main__":
    main()
