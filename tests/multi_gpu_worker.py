"""Worker of tests/test_gpu_multi.py (launched by torchrun, one rank per GPU, NCCL): the N-GPU path on hardware.
Every rank solves its block of the batch device-resident, the results are all-gathered ON THE DEVICE over NCCL, and
rank 0 checks the gathered arrays bit for bit against a single-rank solve of the whole batch and against the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cascade_b200 import api, sharding  # noqa: E402
from cascade_b200.format import entry_major_to_instance_major, random_balls, unpack_entries  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oracle import cbo
    from tests.cases import bessel_shifted
    ok = True
    for L, batch, n in ((32, 1001, 60), (64, 45, 30)):  # odd batches: uneven shards
        rng = np.random.default_rng(99 + L)
        ode = bessel_shifted(cbo, rng, L)
        init = random_balls(rng, batch * 4, L, -1, 0, "zero")
        gm, gt = sharding.solve_fuchs_sharded_device(ode, init, 2, 2, n, batch, True, device=dev)
        assert gm.is_cuda and gm.shape == (n + 1, L, 2 * batch)
        lm, lt = sharding.solve_fuchs_sharded_device(ode, init, 2, 2, n, batch, True, device=dev, gather_rows=[n])
        got = entry_major_to_instance_major(unpack_entries(L, gm.cpu().numpy().view(np.uint32), gt.cpu().numpy()), batch, n + 1)
        last = unpack_entries(L, lm.cpu().numpy().view(np.uint32), lt.cpu().numpy())
        # the host-staged gather of the same sharded solve (any backend) must agree as well
        host = sharding.solve_fuchs_sharded(ode, init, 2, 2, n, batch, True, driver=api.HostBufferDriver(local))
        if rank == 0:
            single = api.acb_ode_solve_fuchs_batch(ode, init, 2, 2, n, batch, True, driver=api.HostBufferDriver(local))
            rc, ref = cbo.fuchs_batch(2, 2, n, ode, init, batch, True, nthreads=8)
            lastref = single.mant.reshape(batch, n + 1, 2, L)[:, n].reshape(-1, L)
            ok = ok and rc == 0 and got.same_as(single) and got.same_as(ref) and host.same_as(single) \
                and np.array_equal(last.mant, lastref)
    # per-instance integer operators (the Legendre sweep of BASELINE configs[3]) sharded the same way
    L, batch, n = 32, 777, 50
    ode = api.acb_ode_legendre_batch(range(batch), L)
    from cascade_b200.format import from_complex
    init = from_complex([1, 1] * batch, L)
    host = sharding.solve_fuchs_sharded(ode, init, 2, 2, n, batch, False, driver=api.HostBufferDriver(local))
    gm, gt = sharding.solve_fuchs_sharded_device(ode, init, 2, 2, n, batch, False, device=dev)
    got = entry_major_to_instance_major(unpack_entries(L, gm.cpu().numpy().view(np.uint32), gt.cpu().numpy()), batch, n + 1)
    if rank == 0:
        single = api.acb_ode_solve_fuchs_batch(ode, init, 2, 2, n, batch, False, driver=api.HostBufferDriver(local))
        ok = ok and host.same_as(single) and got.same_as(single)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK" if ok else "MULTI_GPU_MISMATCH", world, flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
