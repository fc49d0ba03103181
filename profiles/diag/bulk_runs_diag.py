"""Diagnostic for the opt-in bulk-copy RUNS kernel (run with PTG_RUNS_IMPL=b): deposit particles whose
cell encodes their index, so that every lost / duplicated particle can be named (tile, lane, position).
profiles/r01/bulk_runs_diag_before_fix.log / _after_fix.log are its output around the slot hand-back fix;
tests/test_gpu_parity.py runs `run(1 << 21, 1)` as a regression test."""
import sys
import numpy as np
import torch
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from postpic_b200 import device

GW = 148 * 16


def coded(n, per_slice, seed):
    """particle i sits in the centre of a private cell block (blocks 4 cells apart so that TSC stencils
    never overlap); `per_slice` consecutive particles share a block"""
    slices = n // per_slice
    side_a = 1 << ((slices.bit_length() - 1 + 1) // 2)
    side_b = slices // side_a
    perm = np.random.default_rng(seed).permutation(slices)
    cell = np.repeat(perm, per_slice)
    a, b = cell // side_b, cell % side_b
    bins = (4 * side_a, 4 * side_b)
    x = (4 * a + 1 + 0.5) / bins[0]
    y = (4 * b + 1 + 0.5) / bins[1]
    return x, y, bins, perm, side_b


def run(n, per_slice, seed=1):
    x, y, bins, perm, side_b = coded(n, per_slice, seed)
    tx, ty = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    tw = torch.ones(n, dtype=torch.float64, device='cuda')
    kw = dict(range=[(0, 1), (0, 1)], bins=bins, shape=2)
    got, _ = device.deposit([tx, ty], tw, method='runs', **kw)
    a, b = perm // side_b, perm % side_b
    centre = got[torch.from_numpy(4 * a + 1 + 2).cuda(), torch.from_numpy(4 * b + 1 + 2).cuda()].cpu().numpy()
    count = centre / 0.5625
    bad = np.nonzero(np.abs(count - per_slice) > 1e-9)[0]
    total = float(got.sum())
    n_tiles = n // 256
    per_warp = -(-n_tiles // GW)
    print('n=%d per_slice=%d  sum-n=%.3f  bad slices=%d of %d  (tiles/warp=%d)' % (n, per_slice, total - n, len(bad), len(perm), per_warp))
    rows = []
    for s in bad[:4000]:
        first = s * per_slice                       # first particle of the slice
        tile, within = divmod(first, 256)
        lane, j = divmod(within, 8)
        rows.append((tile % per_warp, lane, j, count[s], tile))
    if rows:
        r = np.array(rows)
        print('   tile ordinal within warp: ', dict(zip(*np.unique(r[:, 0].astype(int), return_counts=True))))
        print('   lane:                     ', dict(zip(*np.unique(r[:, 1].astype(int), return_counts=True))))
        print('   position in lane:         ', dict(zip(*np.unique(r[:, 2].astype(int), return_counts=True))))
        print('   deposited count:          ', dict(zip(*np.unique(np.round(r[:, 3], 3), return_counts=True))))
        print('   first few (ordinal, lane, j, count, tile):', rows[:12])
    return len(bad)


if __name__ == '__main__':
    for n, per in ((1 << 16, 8), (1 << 19, 8), (1 << 21, 8), (1 << 22, 8), (1 << 19, 1), (1 << 21, 1), (1 << 22, 1)):
        run(n, per)
        torch.cuda.synchronize()
