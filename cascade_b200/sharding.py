"""Batch partitioning across the GPUs of one box.

Every ODE instance is independent, so the N-GPU path is a contiguous block partition of the batch
index with NO data-path collective; the only communication is an optional final gather of
results (``torch.distributed`` -- NCCL over NVLink on the GPU box, gloo in the CPU tests).
Because the partition is deterministic and no arithmetic crosses ranks, N-rank results are
bit-identical to the single-rank result.
"""
from __future__ import annotations

import numpy as np

from .format import Balls


def partition(batch: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of the instances rank ``rank`` owns: contiguous blocks, sizes differ by at most 1."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_instances(b: Balls, batch: int, per_inst: int, lo: int, hi: int) -> Balls:
    """Rows lo..hi of a ``[batch][per_inst]`` complex-ball array."""
    step = per_inst * 2
    return Balls(b.L, b.mant[lo * step:hi * step].copy(), b.meta[lo * step:hi * step].copy())


def gather_instances(local: Balls, batch: int, per_inst: int, group=None) -> Balls:
    """All-gather of per-rank ``[n_local][per_inst]`` arrays into the full ``[batch][per_inst]`` one
    (uneven shards are padded to the largest and trimmed).  Works on any torch.distributed backend."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    L, step = local.L, per_inst * 2
    nmax = max(partition(batch, world, r)[1] - partition(batch, world, r)[0] for r in range(world))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    m = torch.zeros((nmax * step, L), dtype=torch.int32)
    t = torch.zeros((nmax * step, 4), dtype=torch.int32)
    m[: local.n_slots] = torch.from_numpy(local.mant.view(np.int32))
    t[: local.n_slots] = torch.from_numpy(local.meta)
    m, t = m.to(dev), t.to(dev)
    gm = [torch.empty_like(m) for _ in range(world)]
    gt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gm, m, group=group)
    dist.all_gather(gt, t, group=group)
    parts = []
    for r in range(world):
        lo, hi = partition(batch, world, r)
        n = (hi - lo) * step
        parts.append(Balls(L, gm[r][:n].cpu().numpy().view(np.uint32).copy(), gt[r][:n].cpu().numpy().copy()))
    return Balls.concat(parts)


def solve_fuchs_sharded(ode: Balls, init: Balls, order: int, degree: int, n: int, batch: int,
                        shared_ode: bool = True, driver=None, group=None, gather: bool = True):
    """Each rank solves its block of the batch; optionally all-gather the series."""
    import torch.distributed as dist

    from . import api
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = partition(batch, world, rank)
    # one valuation for the WHOLE batch (checked over every instance, not just this rank's block), so that all
    # ranks agree and a mixed batch is rejected everywhere
    v = api.batch_valuation(ode, order, degree, shared_ode, "solve_fuchs_sharded")
    ne = (order + 1) * (degree + 1)
    my_ode = ode if shared_ode else shard_instances(ode, batch, ne, lo, hi)
    my_init = shard_instances(init, batch, -v, lo, hi)
    local = api.acb_ode_solve_fuchs_batch(my_ode, my_init, order, degree, n, hi - lo, shared_ode, valuation=v,
                                          driver=driver)
    if not gather:
        return local
    return gather_instances(local, batch, n + 1, group=group)


def solve_frobenius_sharded(ode: Balls, rho: Balls, mul: int, alpha: int, order: int, degree: int, n: int, batch: int,
                            shared_ode: bool = False, shared_rho: bool = False, driver=None, group=None,
                            gather: bool = True):
    """Parameter sweep of ``acb_ode_solve_frobenius`` (BASELINE configs[2]): each rank solves its block of the
    (operator, rho) instances, any ``M = mul + alpha``; optionally all-gather ``sol->gens`` (``[batch][M][n+1]``)."""
    import torch.distributed as dist

    from . import api
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = partition(batch, world, rank)
    ne = (order + 1) * (degree + 1)
    v = api.batch_valuation(ode, order, degree, shared_ode, "solve_frobenius_sharded")
    my_ode = ode if shared_ode else shard_instances(ode, batch, ne, lo, hi)
    my_rho = rho if shared_rho else shard_instances(rho, batch, 1, lo, hi)
    local = api.acb_ode_solve_frobenius_batch(my_ode, my_rho, mul, alpha, order, degree, n, hi - lo, shared_ode,
                                              shared_rho, valuation=v, driver=driver)
    if not gather:
        return local
    return gather_instances(local, batch, (mul + alpha) * (n + 1), group=group)


def evaluate_sharded(gens: Balls, M: int, rho: Balls, pts: Balls, driver=None, group=None, gather: bool = True):
    """``acb_ode_solution_evaluate`` of one solution at many points (BASELINE configs[4]): the points are
    partitioned, the series are replicated (a few MB); optionally all-gather the values."""
    import torch.distributed as dist

    from . import api
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    npts = pts.n_slots // 2
    lo, hi = partition(npts, world, rank)
    local = api.acb_ode_solution_evaluate_batch(gens, M, rho, shard_instances(pts, npts, 1, lo, hi), driver=driver)
    if not gather:
        return local
    return gather_instances(local, npts, 1, group=group)


# ---------------------------------------------------------------- device-resident N-GPU path (NCCL over NVLink)
def gather_device(m_local, t_local, batch: int, group=None):
    """All-gather of per-rank DEVICE arrays in the device layout -- ``m_local``: int32 ``[entries][L][2*n_local]``,
    ``t_local``: int32 ``[entries][2*n_local][4]`` -- into ``[entries][L][2*batch]`` / ``[entries][2*batch][4]`` on every
    rank, without touching the host (uneven shards are padded to the largest and trimmed on the device).  NCCL on the
    GPU box; any torch.distributed backend that supports all_gather_into_tensor on these tensors works."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    E, L, _ = m_local.shape
    spans = [partition(batch, world, r) for r in range(world)]
    nmax = max(hi - lo for lo, hi in spans)
    pm = torch.zeros((E, L, 2 * nmax), dtype=m_local.dtype, device=m_local.device)
    pt = torch.zeros((E, 2 * nmax, 4), dtype=t_local.dtype, device=t_local.device)
    pm[:, :, : m_local.shape[2]] = m_local
    pt[:, : t_local.shape[1]] = t_local
    gm = torch.empty((world,) + tuple(pm.shape), dtype=pm.dtype, device=pm.device)
    gt = torch.empty((world,) + tuple(pt.shape), dtype=pt.dtype, device=pt.device)
    # flat views: every backend accepts (world * numel) <- (numel)
    dist.all_gather_into_tensor(gm.view(-1), pm.view(-1), group=group)
    dist.all_gather_into_tensor(gt.view(-1), pt.view(-1), group=group)
    m = torch.cat([gm[r, :, :, : 2 * (hi - lo)] for r, (lo, hi) in enumerate(spans)], dim=2)
    t = torch.cat([gt[r, :, : 2 * (hi - lo)] for r, (lo, hi) in enumerate(spans)], dim=1)
    return m.contiguous(), t.contiguous()


def solve_fuchs_sharded_device(ode: Balls, init: Balls, order: int, degree: int, n: int, batch: int,
                               shared_ode: bool = True, group=None, device=None, gather_rows=None):
    """The N-GPU path as it runs on the box: every rank uploads ITS block of the batch, solves it with the
    device-resident entry point (``cb200_fuchs_batch_dev`` on torch-owned device memory) and the results are
    all-gathered on the device (NCCL).  ``gather_rows``: None = every coefficient, or a list of coefficient indices
    (BASELINE configs[3] gathers a reduction: the last coefficient).  Returns device tensors
    (mant ``[rows][L][2*batch]``, meta ``[rows][2*batch][4]``) in the device layout."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from . import _lib, api
    from .format import instance_major_to_entry_major, pack_entries
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    lo, hi = partition(batch, world, rank)
    nl = hi - lo
    v = api.batch_valuation(ode, order, degree, shared_ode, "solve_fuchs_sharded_device")
    if v >= 0:
        _lib.check(_lib.EVALUATION, "solve_fuchs_sharded_device")
    ne, L = (order + 1) * (degree + 1), ode.L
    my_ode = ode if shared_ode else shard_instances(ode, batch, ne, lo, hi)
    my_init = shard_instances(init, batch, -v, lo, hi)
    ode_m, ode_t = api._ode_to_device(my_ode, order, degree, nl, shared_ode)
    init_m, init_t = pack_entries(instance_major_to_entry_major(my_init, nl, -v), -v, 2 * nl)
    lib = _lib.lib()
    d = lambda a: torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).to(dev)
    d_ode_m, d_ode_t = d(ode_m), d(ode_t)
    res_m = torch.zeros((n + 1, L, 2 * nl), dtype=torch.int32, device=dev)
    res_t = torch.zeros((n + 1, 2 * nl, 4), dtype=torch.int32, device=dev)
    res_m[: -v] = d(init_m)
    res_t[: -v] = d(init_t)
    wsb = lib.cb200_fuchs_workspace_bytes(L, order, degree, v, n, nl, int(shared_ode))
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    if nl > 0:
        _lib.check(lib.cb200_fuchs_batch_dev(L, order, degree, v, n, nl, p(d_ode_m), p(d_ode_t), int(shared_ode), p(res_m),
                                             p(res_t), p(ws), wsb, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   "solve_fuchs_sharded_device")
    if gather_rows is not None:
        idx = torch.tensor(list(gather_rows), device=dev)
        res_m, res_t = res_m.index_select(0, idx), res_t.index_select(0, idx)
    return gather_device(res_m, res_t, batch, group=group)
