"""Multi-GPU deposition: particles shard, grids sum.

Deposition is linear in the particle set, so R ranks (one process per GPU) each
deposit their contiguous block of particles into a full private ghost-padded grid
with identical lo/inv/bins, and one sum over the ranks (NCCL all-reduce or reduce
over NVLink 5 / NVSwitch) yields the reference's result.  There is exactly one
exchange step and it follows the deposit, so the collective is a plain
`torch.distributed` call on the padded grid (8.45 MB for 1028^2, 137 MB for 258^3).
When a range end is missing (`range=None`, _routines.py:115-127) the ranks first
agree on the global min/max (2 doubles per axis).

The reference has no distributed path at all; this module is the B200-native
answer to SURVEY.md section 8(e).
"""
import builtins as _builtins      # `range` is a keyword argument below, as in the reference API

import numpy as np

from . import device as _device
from ._setup import as_bins


def shard_bounds(n, world_size, rank):
    """Contiguous block [begin, end) of rank `rank` when n particles are split over
    `world_size` ranks (the first n % world_size ranks get one extra)."""
    base, extra = divmod(int(n), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def global_range(local_minmax, group=None):
    """All-reduce per-axis (min, max) pairs so that every rank builds identical edges.

    local_minmax: list of (lo, hi) Python floats (one pair per axis), +inf/-inf for a
    rank without particles.  Works on any backend (tensors are placed on the GPU for
    NCCL, on the CPU for gloo)."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    lo = torch.tensor([p[0] for p in local_minmax], dtype=torch.float64, device=dev)
    hi = torch.tensor([p[1] for p in local_minmax], dtype=torch.float64, device=dev)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return list(zip(lo.tolist(), hi.tolist()))


def sum_grids(grid, group=None, dst=None):
    """In-place sum of the ranks' padded grids.  dst=None: all-reduce (every rank ends
    up with the field); dst=k: reduce to rank k only (half the NVLink traffic, enough
    when one rank hands the array back to the host)."""
    import torch.distributed as dist
    if dst is None:
        dist.all_reduce(grid, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(grid, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return grid


def deposit_sharded(coords, weights=None, range=None, bins=None, shape=2, method=None, group=None,
                    dst=None, local_deposit=None):
    """Deposit this rank's shard and sum over ranks.

    coords/weights: this rank's particles (CUDA tensors).  range entries may be None
    (or the whole range None), in which case the global min/max is used, as
    histogramdd does on one process.  Returns (padded grid tensor, plan); slice with
    `plan.interior`.  `local_deposit(grid, plan, coords, weights)` can replace the CUDA
    deposit -- the CPU tests inject the oracle there to exercise this host logic
    under gloo."""
    import torch
    coords = list(coords)
    nd = len(coords)
    rng = [[None, None] for _ in coords]
    if range is not None:
        r = [range] if (nd == 1 and np.isscalar(range[0])) else range
        for a in _builtins.range(nd):
            if r[a] is not None:
                rng[a] = [r[a][0], r[a][1]]
    missing = [a for a in _builtins.range(nd) if rng[a][0] is None or rng[a][1] is None]
    if missing:
        local = []
        for a in missing:
            if coords[a].numel() == 0:
                local.append((float('inf'), float('-inf')))
            elif coords[a].is_cuda:
                local.append(_device.minmax(coords[a]))
            else:
                local.append((float(coords[a].min()), float(coords[a].max())))
        glob = global_range(local, group)
        for a, (lo, hi) in zip(missing, glob):
            rng[a][0] = lo if rng[a][0] is None else rng[a][0]
            rng[a][1] = hi if rng[a][1] is None else rng[a][1]
    b = as_bins(bins[0] if (nd == 1 and np.ndim(bins)) else bins, nd)
    plan = _device.Plan(rng, b, shape)
    grid = torch.empty(plan.dims, dtype=torch.float64, device=coords[0].device)
    if local_deposit is None:
        _device.deposit_into(grid, plan, coords, weights, method=method, zero=True)
    else:
        local_deposit(grid, plan, coords, weights)
    sum_grids(grid, group=group, dst=dst)
    return grid, plan



def gpu_numa_node(device_index):
    """NUMA node of GPU `device_index`'s PCIe root as sysfs reports it (None when unknown)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        with open("/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:].lower(), rest.lower())) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def prefer_numa_node(node):
    """set_mempolicy(MPOL_PREFERRED, {node}) for this thread (inherited by the threads it starts): pages this process
    allocates from now on -- the pinned buffers above all -- come from `node` whichever CPU the thread runs on (the CPU
    binding alone decides nothing when the container's cpuset holds no CPU of that node).  Returns True on success."""
    import ctypes
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238           # x86_64
        return libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(16 * 64 + 1)) == 0
    except Exception:
        return False


def numa_pages_of(address):
    """{node: pages} of the mapping that holds `address`, from /proc/self/numa_maps (diagnostics: where did a pinned
    buffer land); None when the file cannot be read."""
    try:
        best = None
        with open("/proc/self/numa_maps") as f:
            for line in f:
                parts = line.split()
                start = int(parts[0], 16)
                if start <= address and (best is None or start > best[0]):
                    best = (start, parts)
        if best is None:
            return None
        return {int(p[1:].split("=")[0]): int(p.split("=")[1]) for p in best[1] if p[0] == "N" and "=" in p and p[1].isdigit()}
    except Exception:
        return None


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs NVML reports as local to GPU `device_index` and prefer the memory of the GPU's NUMA
    node, so that the pinned host buffers it allocates afterwards and the staging threads of the host pipeline sit on
    the NUMA node the GPU's PCIe root hangs off.  Returns the CPU list, or None when NVML / the affinity call is not
    available or the process may not run on any of those CPUs (the memory preference is still tried then)."""
    import os
    node = gpu_numa_node(device_index)
    if node is not None and os.environ.get("PTG_NO_MEMPOLICY") is None:
        prefer_numa_node(node)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * k + b for k, wd in enumerate(mask) for b in _builtins.range(64) if (wd >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def histogram_host_sharded(coords, weights=None, range=None, bins=None, shape=2, method=None, group=None,
                           dst=0, device_index=None, local_deposit=None, local_minmax=None):
    """The numpy drop-in over R ranks: every rank passes ITS block of the particles as host numpy arrays.

    Each rank streams its block through the library's chunked H2D pipeline into a private padded grid in
    HBM (ptg_deposit_host_f64: the PCIe copy of chunk k+1 overlaps the deposit of chunk k), the grids are
    summed with one NCCL reduce to rank `dst` over NVLink (dst=None: all-reduce), and only that rank copies
    the sum back to the host.  Returns (padded numpy grid or None on the other ranks, plan).  Missing range
    ends are completed with the global min / max like histogramdd does on one process.
    `local_deposit(grid, plan, coords, weights)` / `local_minmax(array) -> (lo, hi)` replace the CUDA steps (the
    grid is then a CPU tensor): the gloo tests inject the oracle there to exercise this host logic without a GPU."""
    import ctypes
    import torch
    import torch.distributed as dist
    from . import _lib
    coords = [np.asarray(c) for c in coords]
    nd = len(coords)
    if device_index is None and local_deposit is None:
        device_index = torch.cuda.current_device()
    rng = [[None, None] for _ in coords]
    if range is not None:
        r = [range] if (nd == 1 and np.isscalar(range[0])) else range
        for a in _builtins.range(nd):
            if r[a] is not None:
                rng[a] = [r[a][0], r[a][1]]
    missing = [a for a in _builtins.range(nd) if rng[a][0] is None or rng[a][1] is None]
    if missing:
        local = []
        for a in missing:
            c = coords[a]
            if c.shape[0] == 0:
                local.append((float('inf'), float('-inf')))
                continue
            if local_minmax is not None:
                local.append(tuple(local_minmax(c)))
                continue
            out = (ctypes.c_double * 2)()
            _lib.check(_lib.load().ptg_minmax_host_f64(c.shape[0], c.ctypes.data, c.strides[0] // 8 if c.shape[0] > 1 else 1,
                                                       out, int(device_index)))
            local.append((out[0], out[1]))
        for a, (lo, hi) in zip(missing, global_range(local, group)):
            rng[a][0] = lo if rng[a][0] is None else rng[a][0]
            rng[a][1] = hi if rng[a][1] is None else rng[a][1]
    b = as_bins(bins[0] if (nd == 1 and np.ndim(bins)) else bins, nd)
    plan = _device.Plan(rng, b, shape)
    if local_deposit is None:
        grid = torch.empty(plan.dims, dtype=torch.float64, device=torch.device('cuda', int(device_index)))
        _device.deposit_host_into(grid, plan, coords, weights, method=method, zero=True)
    else:
        grid = torch.zeros(plan.dims, dtype=torch.float64)
        local_deposit(grid, plan, coords, weights)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        sum_grids(grid, group=group, dst=dst)
        if dst is not None and dist.get_rank(group) != dst:
            return None, plan
    return grid.cpu().numpy(), plan
