"""BASELINE configs[1] at FULL size on the GPU (65 536 instances x 2 001 coefficients, 1024 bits,
device resident through the C ABI), checked three ways: a sample of instances bit for bit against
the oracle, the reference's own acceptance property (residual contains zero) on sampled instances,
and finiteness / monotone radius growth over the whole 37.7 GB result."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_full_config_sampled_parity(oracle):
    import torch

    import bench
    from cascade_b200 import _lib, api
    from cascade_b200.format import (Balls, FLAG_NAN, entry_major_to_instance_major, instance_major_to_entry_major,
                                     pack_entries, unpack_entries)

    lib = _lib.lib()
    assert lib.cb200_device_count() > 0
    L, n, batch = bench.L, bench.N_TERMS, bench.BATCH
    dev = torch.device("cuda:0")
    ode, init = bench.make_inputs(batch, n, api.HostBufferDriver(0))
    S = 2 * batch
    ode_m, ode_t = api._ode_to_device(ode, 2, 2, batch, True)
    init_m, init_t = pack_entries(instance_major_to_entry_major(init, batch, 2), 2, S)
    d_ode_m = torch.from_numpy(ode_m.view(np.int32)).to(dev)
    d_ode_t = torch.from_numpy(ode_t).to(dev)
    d_res_m = torch.empty((n + 1, L, S), dtype=torch.int32, device=dev)
    d_res_t = torch.empty((n + 1, S, 4), dtype=torch.int32, device=dev)
    d_res_m[:2].copy_(torch.from_numpy(init_m.view(np.int32)))
    d_res_t[:2].copy_(torch.from_numpy(init_t))
    wsb = lib.cb200_fuchs_workspace_bytes(L, 2, 2, -2, n, batch, 1)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.cb200_fuchs_batch_dev(L, 2, 2, -2, n, batch, p(d_ode_m), p(d_ode_t), 1, p(d_res_m), p(d_res_t), p(ws), wsb, None)
    assert rc == 0
    torch.cuda.synchronize()
    # whole result: no indeterminate ball anywhere; radii are nonzero from the first rounded step on
    flags = d_res_t[:, :, 1]
    assert int((flags & FLAG_NAN).sum().item()) == 0
    assert int((d_res_t[4:, :, 2] == 0).sum().item()) == 0
    # sampled instances: first/last of warps and blocks, and random ones
    rng = np.random.default_rng(7)
    sample = sorted({0, 1, 31, 32, 223, 224, 225, 33151, 33152, batch - 1, *rng.integers(0, batch, 6).tolist()})
    idx = torch.tensor([2 * i + part for i in sample for part in (0, 1)], device=dev)
    m = d_res_m.index_select(2, idx).cpu().numpy().view(np.uint32)   # [n+1, L, 2*len]
    t = d_res_t.index_select(1, idx).cpu().numpy()
    got = entry_major_to_instance_major(unpack_entries(L, m, t), len(sample), n + 1)
    sub_init = Balls(L, np.concatenate([init.mant[4 * i:4 * i + 4] for i in sample]),
                     np.concatenate([init.meta[4 * i:4 * i + 4] for i in sample]))
    rc, want = oracle.fuchs_batch(2, 2, n, ode, sub_init, len(sample), True, nthreads=8)
    assert rc == 0
    assert got.same_as(want), "full-size GPU run differs from the oracle on the sampled instances"
    # the reference's acceptance test (src/acb_ode.c:208-240) on two of them
    for k in (0, len(sample) - 1):
        series = got[k * (n + 1) * 2:(k + 1) * (n + 1) * 2]
        assert oracle.ode_solves(2, 2, ode, series, 200)


@pytest.mark.parametrize("persistent", ["0", "1"])
def test_coefficient_ranges_match_one_call(monkeypatch, persistent):
    """cb200_fuchs_batch_range_dev: solving coefficient ranges one after the other (what bench.py's end-to-end
    leg does to overlap the copy-out) leaves exactly the bytes of a single full call -- with whole-wave launches
    and with the persistent kernel (whose work units then start at the range's first coefficient)."""
    import torch

    monkeypatch.setenv("CB200_PERSISTENT", persistent)

    import bench
    from cascade_b200 import _lib, api
    from cascade_b200.format import instance_major_to_entry_major, pack_entries

    lib = _lib.lib()
    L, n, batch = bench.L, 300, 4096
    dev = torch.device("cuda:0")
    ode, init = bench.make_inputs(batch, n, api.HostBufferDriver(0))
    S = 2 * batch
    ode_m, ode_t = api._ode_to_device(ode, 2, 2, batch, True)
    init_m, init_t = pack_entries(instance_major_to_entry_major(init, batch, 2), 2, S)
    d_ode_m = torch.from_numpy(ode_m.view(np.int32)).to(dev)
    d_ode_t = torch.from_numpy(ode_t).to(dev)
    wsb = lib.cb200_fuchs_workspace_bytes(L, 2, 2, -2, n, batch, 1)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    outs = []
    for ranges in ([(0, n)], [(0, 50), (51, 51), (52, 199), (200, n)]):
        d_res_m = torch.zeros((n + 1, L, S), dtype=torch.int32, device=dev)
        d_res_t = torch.zeros((n + 1, S, 4), dtype=torch.int32, device=dev)
        d_res_m[:2].copy_(torch.from_numpy(init_m.view(np.int32)))
        d_res_t[:2].copy_(torch.from_numpy(init_t))
        for lo, hi in ranges:
            rc = lib.cb200_fuchs_batch_range_dev(L, 2, 2, -2, lo, hi, batch, p(d_ode_m), p(d_ode_t), 1, p(d_res_m),
                                                 p(d_res_t), p(ws), wsb, None)
            assert rc == 0
        torch.cuda.synchronize()
        outs.append((d_res_m.cpu(), d_res_t.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("L,batch", [(64, 1000), (32, 1500)])
def test_persistent_kernel_matches_whole_wave_kernel(monkeypatch, L, batch):
    """The work-unit (persistent) recurrence kernel and the one-block-per-instance-block kernel leave the same
    bytes, ragged last block included -- at 2048 bits too, where the persistent form serves batches that are not
    a whole number of waves (the continuation pipeline's)."""
    from cascade_b200 import api
    from cascade_b200.format import random_balls

    rng = np.random.default_rng(0xBE55 + L)
    drv = api.HostBufferDriver(0)
    nu = random_balls(rng, 2, L, 0, 0, "zero")
    a = random_balls(rng, 2, L, 0, 1, "zero")
    ode0, _ = api.acb_ode_bessel_batch(nu, driver=drv)
    ode = api.acb_ode_shift_batch(ode0, a, 2, 2, driver=drv)
    init = random_balls(rng, batch * 4, L, -1, 0, "ulp")
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("CB200_PERSISTENT", flag)
        outs.append(api.acb_ode_solve_fuchs_batch(ode, init, 2, 2, 150, batch, True, driver=drv))
    assert outs[0].same_as(outs[1])
