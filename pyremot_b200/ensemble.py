"""Ensemble sharding across the GPUs of one box: one process per GPU, no data-path collective.

Reactor instances are independent, so rank r simply owns a slice of the instance table and the only
communication is ONE gather of outlet rows at the end (NCCL over NVLink when the tensors are CUDA tensors;
gloo in the CPU tests).  ``torch.distributed`` is plumbing here; everything numerical happens in libpbr.

Load balance: BDF step counts differ between easy and hot-spot reactors, so a contiguous split of a table that
is ordered by operating point would give ranks unequal work.  ``deal()`` sorts by a stiffness proxy and deals
the rows round-robin; ``undeal()`` restores the caller's order after the gather.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(N, rank, world):
    """Contiguous shard [lo, hi) of rank; the first N % world ranks get one extra row."""
    q, r = divmod(N, world)
    lo = rank*q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def deal(proxy, world):
    """Permutation that sorts instances by ``proxy`` and deals them round-robin to ``world`` ranks.

    Returns ``perm`` such that rank r owns rows ``perm[lo_r:hi_r]`` (with shard_bounds); every rank gets the
    same mix of easy and hard instances."""
    N = len(proxy)
    order = np.argsort(proxy, kind="stable")
    perm = np.empty(N, dtype=np.int64)
    pos = 0
    for r in range(world):
        mine = order[r::world]
        perm[pos:pos + len(mine)] = mine
        pos += len(mine)
    return perm


def undeal(perm, gathered):
    """Inverse of ``deal`` on the gathered rows: out[perm[k]] = gathered[k]."""
    out = np.empty_like(gathered)
    out[perm] = gathered
    return out


def gather_rows(local, N_total, world=None, rank=None):
    """All-gather the per-rank row blocks (torch tensor [n_r, K]) into [N_total, K] on every rank.

    Shards may differ by one row: blocks are padded to the largest shard for the collective and trimmed after."""
    import torch
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank() if rank is None else rank
    K = local.shape[1]
    nmax = -(-N_total // world)
    pad = torch.zeros((nmax, K), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world*nmax, K), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(N_total, r, world)
        parts.append(out[r*nmax:r*nmax + (hi - lo)])
    return torch.cat(parts, dim=0)


def solve_n1_sharded(modelInput, sweep, params=(), rtol=None, atol=None, balance=True, device=None):
    """Solve an N1 ensemble across all ranks of the initialised process group; every rank returns the full
    outlet table [N, nv1] (numpy) in the caller's row order, plus its local status vector."""
    import torch
    import torch.distributed as dist
    from .api import rmtExeBatch
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    N = len(next(iter(sweep.values())))
    perm = deal(np.asarray(sweep["T"]) if (balance and "T" in sweep) else np.arange(N), world) if world > 1 else np.arange(N)
    lo, hi = shard_bounds(N, rank, world)
    mine = perm[lo:hi]
    local_sweep = {k: np.asarray(v)[mine] for k, v in sweep.items()}
    dev = torch.cuda.current_device() if device is None else device
    if hi > lo:
        r = rmtExeBatch(modelInput, local_sweep, params=params, device=dev, rtol=rtol, atol=atol, return_device=True)
        outlet, status = r["outlet"].contiguous(), r["status"].cpu().numpy()
    else:                                   # empty shard (fewer reactors than ranks): nothing to solve, still take part in the gather
        from .api import get_model
        nv1 = get_model(modelInput, params=params, device=dev).nv1
        outlet = torch.empty((0, nv1), dtype=torch.float64, device=torch.device("cuda", dev))
        status = np.empty((0,), dtype=np.int32)
    full = gather_rows(outlet, N, world, rank)
    return undeal(perm, full.cpu().numpy()), status


def chunk_bounds(N, chunk):
    """[(a, b), ...] covering [0, N) in pieces of at most ``chunk`` rows."""
    chunk = max(1, int(chunk))
    return [(a, min(a + chunk, N)) for a in range(0, N, chunk)]


def solve_n2_sharded(modelInput, sweep, zNo=None, tNo=None, chunk=16384, rtol=None, atol=None, device=None, keep=None):
    """Dynamic-model (N2) ensemble across all ranks of the initialised process group (BASELINE config 4: zNo = 400,
    1,048,576 reactors = 131,072 per GPU at 8 GPUs).

    Every rank solves its contiguous shard in chunks of ``chunk`` reactors: the full snapshots of a chunk
    ([chunk, tNo, n+1, zNo] doubles; 1.8 GB at 16,384 x zNo = 400) exist on the GPU only while the chunk is processed.
    Kept per reactor: the outlet-node rows ``[x_i ; T]`` of every snapshot ([tNo*(n+1)]), gathered with ONE all-gather at
    the end; full profiles stay with the rank that made them -- ``keep(lo, hi, dataYs_chunk, status_chunk)`` is called
    with each chunk's device tensors (global row range) so the caller can write them out (``io.save_shard``).
    Returns (outlet rows [N, tNo, n+1] numpy on every rank, local status [n_local] numpy, mean BDF steps of the shard).
    """
    import torch
    import torch.distributed as dist
    from .api import DEFAULT_ATOL, DEFAULT_RTOL, get_model, solverSetting
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    s2 = solverSetting["N2"]
    zNo = int(zNo or s2["zNo"])
    tNo = int(tNo or s2["tNo"])
    rtol = DEFAULT_RTOL if rtol is None else rtol
    atol = DEFAULT_ATOL if atol is None else atol
    N = len(next(iter(sweep.values())))
    lo, hi = shard_bounds(N, rank, world)
    dev = torch.cuda.current_device() if device is None else device
    model = get_model(modelInput, device=dev)
    tab = model.instance_table(modelInput, {k: np.asarray(v)[lo:hi] for k, v in sweep.items()})
    tab = torch.from_numpy(tab).to(model._dev())
    par = model.prolog(tab)
    n = model.n
    period = modelInput["operating-conditions"]["period"]
    outlet = torch.empty((hi - lo, tNo*(n + 1)), dtype=torch.float64, device=tab.device)
    status = torch.empty((hi - lo,), dtype=torch.int32, device=tab.device)
    steps = torch.zeros((), dtype=torch.float64, device=tab.device)
    for a, b in chunk_bounds(hi - lo, chunk):
        pc = par[:, a:b].contiguous()
        r = model.solve_n2(pc, zNo, tNo, period, rtol=rtol, atol=atol)
        data = model.unscale_n2(pc, r["snaps_y"])                       # [chunk, tNo, n+1, zNo]
        outlet[a:b] = data[:, :, :, -1].reshape(b - a, -1)
        status[a:b] = r["status"]
        steps += r["stats"][0].double().sum()
        if keep is not None:
            keep(lo + a, lo + b, data, r["status"])
        del r, data
    full = gather_rows(outlet, N, world, rank)
    return full.cpu().numpy().reshape(N, tNo, n + 1), status.cpu().numpy(), float(steps)/max(1, hi - lo)
