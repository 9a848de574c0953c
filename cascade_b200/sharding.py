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
