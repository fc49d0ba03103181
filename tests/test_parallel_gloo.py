"""world_size-2 test of the multi-GPU host logic on CPU (gloo): shard -> private grid
-> sum of grids, and the global min/max agreement for range=None.  The per-rank
deposit is the oracle (injected through `local_deposit`; tests/ may use it)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import ptg_oracle
    from postpic_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(31)
    n = 20001
    x, y, w = rng.uniform(0, 1, n), rng.uniform(-1, 1, n), rng.uniform(0, 1, n)
    b, e = parallel.shard_bounds(n, world, rank)

    def local(grid, plan, coords, weights):
        g, _, _ = ptg_oracle.deposit([c.numpy() for c in coords], None if weights is None else weights.numpy(),
                                     [(plan.edges[a][0], plan.edges[a][-1]) for a in range(plan.ndim)],
                                     plan.bins, plan.shape)
        grid.copy_(torch.from_numpy(g))

    tx, ty, tw = (torch.from_numpy(a[b:e].copy()) for a in (x, y, w))
    results = {}
    # explicit ranges, all-reduce
    g, plan = parallel.deposit_sharded([tx, ty], tw, range=((0, 1), (-1, 1)), bins=(32, 24), shape=2,
                                       local_deposit=local)
    results["allreduce"] = g.numpy().copy()
    # missing ranges: must agree on the global min/max
    g, plan = parallel.deposit_sharded([tx, ty], tw, range=((None, 1.0), None), bins=(32, 24), shape=3,
                                       local_deposit=local)
    results["auto_range"] = g.numpy().copy()
    results["auto_edges0"] = plan.edges[0]
    results["auto_edges1"] = plan.edges[1]
    # reduce to rank 0 only
    g, plan = parallel.deposit_sharded([tx], None, range=(0, 1), bins=50, shape=1, dst=0, local_deposit=local)
    results["reduce0"] = g.numpy().copy()
    # the host-array entry (numpy shards in, reduce to rank 0, only rank 0 gets the grid back)
    def local_np(grid, plan, coords, weights):
        g, _, _ = ptg_oracle.deposit(list(coords), weights, [(plan.edges[a][0], plan.edges[a][-1]) for a in range(plan.ndim)],
                                     plan.bins, plan.shape)
        grid.copy_(torch.from_numpy(g))
    h, plan = parallel.histogram_host_sharded([x[b:e], y[b:e]], w[b:e], range=((0, 1), None), bins=(32, 24), shape=2, dst=0,
                                              local_deposit=local_np, local_minmax=lambda a: (a.min(), a.max()))
    assert (h is None) == (rank != 0)
    if rank == 0:
        results["host_sharded"] = h
        results["host_edges1"] = plan.edges[1]
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), **results)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_deposit_world_size_2(tmp_path, oracle):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    rng = np.random.default_rng(31)
    n = 20001
    x, y, w = rng.uniform(0, 1, n), rng.uniform(-1, 1, n), rng.uniform(0, 1, n)
    full, _, _ = oracle.deposit([x, y], w, [(0, 1), (-1, 1)], [32, 24], 2)
    assert np.array_equal(r0["allreduce"], r1["allreduce"])
    assert np.allclose(r0["allreduce"], full, rtol=1e-13, atol=1e-15)
    full, edges, _ = oracle.deposit([x, y], w, [(x.min(), 1.0), (y.min(), y.max())], [32, 24], 3)
    assert np.array_equal(r0["auto_edges0"], edges[0]) and np.array_equal(r1["auto_edges1"], edges[1])
    assert np.allclose(r0["auto_range"], full, rtol=1e-13, atol=1e-15)
    full, _, _ = oracle.deposit([x], None, [(0, 1)], [50], 1)
    assert np.allclose(r0["reduce0"], full, rtol=1e-13, atol=1e-15)
    full, edges, _ = oracle.deposit([x, y], w, [(0, 1), (y.min(), y.max())], [32, 24], 2)
    assert np.allclose(r0["host_sharded"], full, rtol=1e-13, atol=1e-15) and np.array_equal(r0["host_edges1"], edges[1])
    assert "host_sharded" not in r1
