"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun, NCCL).  Every rank deposits its block of a
seeded cloud; the summed grid is compared with the oracle's histogram of the WHOLE cloud, per bin, on rank 0 --
for the device-resident path (parallel.deposit_sharded) and the host-array path (parallel.histogram_host_sharded),
with ranges given and with ranges completed from the data."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from postpic_b200 import parallel
    from oracle import ptg_oracle
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(2024)                      # every rank draws the same cloud and takes its block
    n = 6_000_000
    x, px, w = rng.random(n), rng.uniform(-1.0, 1.0, n), rng.random(n)
    b, e = parallel.shard_bounds(n, world, rank)
    cases = [dict(range=((0.0, 1.0), (-1.0, 1.0)), bins=(1024, 1024), shape=2),       # the headline geometry
             dict(range=None, bins=(300, 200), shape=3),                                  # ranges agreed on by MIN/MAX all-reduce
             dict(range=((0.1, 0.9), None), bins=(128, 96), shape=1)]
    for kw in cases:
        if rank == 0:
            rg = kw["range"]
            full = [[None, None], [None, None]] if rg is None else [list(r) if r is not None else [None, None] for r in rg]
            for a, arr in enumerate((x, px)):
                if full[a][0] is None:
                    full[a][0] = arr.min()
                if full[a][1] is None:
                    full[a][1] = arr.max()
            h = ptg_oracle.histogram2d(x, px, weights=w, range=full, bins=kw["bins"], shape=kw["shape"])[0]
            ref = h.base if h.base is not None else h
        # device-resident shards, all-reduce
        g, plan = parallel.deposit_sharded([torch.from_numpy(x[b:e]).to(dev), torch.from_numpy(px[b:e]).to(dev)],
                                           torch.from_numpy(w[b:e]).to(dev), **kw)
        # host shards through the chunked pipeline, reduce to rank 0
        hh, plan2 = parallel.histogram_host_sharded([x[b:e], px[b:e]], w[b:e], dst=0, device_index=local, **kw)
        if rank == 0:
            for got, tag in ((g.cpu().numpy(), "deposit_sharded"), (hh, "histogram_host_sharded")):
                assert got.shape == ref.shape, tag
                assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref)), (tag, kw)
                assert abs(got.sum() - ref.sum()) <= 1e-12 * np.abs(w).sum(), (tag, kw)
        else:
            assert hh is None
    # shape 0, integer weights: bit-exact across ranks
    wi = rng.integers(1, 9, n).astype(np.float64)
    kw = dict(range=((0.0, 1.0), (-1.0, 1.0)), bins=(1024, 1024), shape=0)
    g, _ = parallel.deposit_sharded([torch.from_numpy(x[b:e]).to(dev), torch.from_numpy(px[b:e]).to(dev)],
                                    torch.from_numpy(wi[b:e]).to(dev), **kw)
    if rank == 0:
        h = ptg_oracle.histogram2d(x, px, weights=wi, **kw)[0]
        assert np.array_equal(g.cpu().numpy(), h.base if h.base is not None else h)
        print("MULTI_GPU_PARITY_OK world=%d" % world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
