"""world_size-2 run of the N>1 path on CPU (gloo): each rank solves its block of the batch with
the host-simulated device code, results are all-gathered, and must be BIT-IDENTICAL to the
single-rank result (pure batch partitioning)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from cascade_b200 import api, sharding
from cascade_b200.format import random_balls


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cbo
    from tests.cases import bessel_shifted
    from tests.host_sim.driver import SimDriver
    rng = np.random.default_rng(99)
    L, batch, n = 4, 7, 24  # odd batch: uneven shards
    ode = bessel_shifted(cbo, rng, L)
    init = random_balls(rng, batch * 4, L, -1, 0, "zero")
    full = sharding.solve_fuchs_sharded(ode, init, 2, 2, n, batch, True, driver=SimDriver())
    # the device-layout gather (what the NCCL path uses on the box), here on CPU tensors over gloo
    import torch
    from cascade_b200.format import entry_major_to_instance_major, instance_major_to_entry_major, pack_entries, unpack_entries
    lo, hi = sharding.partition(batch, world, rank)
    mine = sharding.shard_instances(full, batch, n + 1, lo, hi)
    lm, lt = pack_entries(instance_major_to_entry_major(mine, hi - lo, n + 1), n + 1, 2 * (hi - lo))
    gm, gt = sharding.gather_device(torch.from_numpy(lm.view(np.int32)), torch.from_numpy(lt), batch)
    regathered = entry_major_to_instance_major(unpack_entries(L, gm.numpy().view(np.uint32), gt.numpy()), batch, n + 1)
    assert regathered.same_as(full), "device-layout gather differs from the instance-major gather"
    # parameter sweep of the (logarithmic) Frobenius solver and a point sweep of the evaluation
    from cascade_b200.format import Balls
    par = random_balls(rng, batch * 6, L, -1, 1, "zero")
    odes = Balls.concat([cbo.example("hypgeom", L, 0, par[i * 6:(i + 1) * 6]) for i in range(batch)])
    rho = random_balls(rng, batch * 2, L, -1, 1, "ulp")
    gens = sharding.solve_frobenius_sharded(odes, rho, 2, 0, 2, 2, 6, batch, driver=SimDriver())
    pts = random_balls(rng, 2 * 5, L, -2, 0, "zero")
    g0 = gens[0:2 * 2 * 7]  # the two series of instance 0
    vals = sharding.evaluate_sharded(g0, 2, rho[0:2], pts, driver=SimDriver())
    if rank == 0:
        single = api.acb_ode_solve_fuchs_batch(ode, init, 2, 2, n, batch, True, driver=SimDriver())
        rc, ref = cbo.fuchs_batch(2, 2, n, ode, init, batch, True)
        gref = cbo.frobenius_general_batch(2, 2, 6, odes, rho, batch, False, False, 2, 2)
        mode, e = api.pow_mode_of(rho[0:2])
        vref = cbo.solution_evaluate(g0, 2, rho[0:2], mode, e, pts)
        out.put((full.same_as(single), full.same_as(ref) and gens.same_as(gref) and vals.same_as(vref)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_bit_identical():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    same_single, same_oracle = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same_single and same_oracle
