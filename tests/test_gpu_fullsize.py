"""BASELINE configs[1] .. configs[4] at FULL size on the GPU (device resident through the C ABI, the kernel form the
LAUNCHER picks at that size -- nothing forced through the environment), each checked three ways: a sample of
instances (first / last of warps and blocks, and random ones) bit for bit against the oracle, the reference's own
acceptance property (residual contains zero) on sampled instances, and finiteness over the whole result."""
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


def _sample(count, edges, seed, extra=6):
    rng = np.random.default_rng(seed)
    return sorted({i for i in edges if 0 <= i < count} | {0, count - 1} | set(rng.integers(0, count, extra).tolist()))


def _pick(dm, dt, sample, L, entries):
    """instances `sample` of device arrays [entries][L][2*count] / [entries][2*count][4] -> host Balls [len(sample)][entries]"""
    import torch

    from cascade_b200.format import entry_major_to_instance_major, unpack_entries
    idx = torch.tensor([2 * i + part for i in sample for part in (0, 1)], device=dm.device)
    m = dm.index_select(2, idx).cpu().numpy().view(np.uint32)
    t = dt.index_select(1, idx).cpu().numpy()
    return entry_major_to_instance_major(unpack_entries(L, m, t), len(sample), entries)


@pytest.mark.parametrize("L,npts", [(64, 20000), (128, 40000)])
def test_horner_cooperative_form_chosen_by_the_launcher_beyond_one_wave(oracle, L, npts):
    """More points than one wave of cooperative CTAs (9 472): the launcher still picks four lanes per point at 2048 bits
    up to three waves and at 4096 bits at every size (phase barriers, ragged last CTA); sampled points against the
    oracle, one and two derivatives."""
    from cascade_b200 import api
    from cascade_b200.format import Balls, random_balls

    rng = np.random.default_rng(4100 + L)
    drv = api.HostBufferDriver(0)
    length = 14
    f = random_balls(rng, 2 * length, L, -3, 3, "ulp")
    pts = random_balls(rng, 2 * npts, L, -2, 0, "zero")
    sample = _sample(npts, [1, 7, 8, 63, 64, 65, 9471, 9472, 9473, npts - 2], 7)
    sub = Balls.concat([pts[2 * i:2 * i + 2] for i in sample])
    for nder in (1, 2):
        got = api.acb_poly_evaluate_batch(f, pts, nder, driver=drv)
        want = oracle.horner_batch(f, sub, nder)
        pick = Balls.concat([got[2 * nder * i:2 * nder * (i + 1)] for i in sample])
        assert pick.same_as(want), f"Horner L={L} nder={nder}: launcher-chosen cooperative kernel differs from the oracle"


def test_config2_both_exponents_in_one_batch_full_size(oracle):
    """BASELINE configs[2] as `bench.py --config 2` submits it: the 2 x 16 384 (operator, exponent) pairs of both runs in
    ONE call (rho = 0 for the first half, rho = 1 - c for the second).  More than one wave of the all-shared kernel, so
    the launcher picks the persistent 208-thread kernel on hybrid columns (work units of 8 coefficients); sampled
    instances at CTA / warp edges of both halves against the oracle."""
    import torch

    import bench
    from cascade_b200 import _lib, api
    from cascade_b200.format import FLAG_NAN, Balls, from_complex, instance_major_to_entry_major, pack_entries

    lib = _lib.lib()
    L, half, n = 128, 16384, 256
    batch = 2 * half
    dev = torch.device("cuda:0")
    drv = api.HostBufferDriver(0)
    a, b, c = bench.hypgeom_params(half, L)
    ode = api.acb_ode_hypgeom_batch(a, b, c, driver=drv)
    rho = Balls.concat([Balls.zeros(2 * half, L), api.ew(api.OP_SUB, from_complex([1] * half, L), c, driver=drv)])
    om, ot = pack_entries(instance_major_to_entry_major(ode, half, 9), 9, 2 * half)
    d_om = torch.from_numpy(np.concatenate([om, om], axis=2).view(np.int32)).to(dev)
    d_ot = torch.from_numpy(np.concatenate([ot, ot], axis=1)).to(dev)
    rm, rt = pack_entries(rho, 1, 2 * batch)
    d_rm, d_rt = torch.from_numpy(rm.view(np.int32)).to(dev), torch.from_numpy(rt).to(dev)
    d_res_m = torch.empty((n + 1, L, 2 * batch), dtype=torch.int32, device=dev)
    d_res_t = torch.empty((n + 1, 2 * batch, 4), dtype=torch.int32, device=dev)
    wsb = lib.cb200_frobenius_workspace_bytes(L, batch)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.cb200_frobenius1_batch_dev(L, 2, 2, -1, n, batch, p(d_om), p(d_ot), 0, p(d_rm), p(d_rt), 0, p(d_res_m), p(d_res_t),
                                        p(ws), wsb, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert int((d_res_t[:, :, 1] & FLAG_NAN).sum().item()) == 0
    sample = _sample(batch, [1, 31, 32, 207, 208, 209, 415, 416, half - 1, half, half + 1, 32655, 32656, batch - 2], 22, extra=4)
    got = _pick(d_res_m, d_res_t, sample, L, n + 1)
    sub_ode = Balls.concat([ode[18 * (i % half):18 * (i % half) + 18] for i in sample])
    sub_rho = Balls.concat([rho[2 * i:2 * i + 2] for i in sample])
    want = oracle.frobenius1_batch(2, 2, n, sub_ode, sub_rho, len(sample), False, nthreads=8)
    assert got.same_as(want), "configs[2], both exponents in one batch: full-size GPU run differs from the oracle on the sampled instances"


def test_frobenius_2048_bits_two_lane_form_chosen_by_the_launcher(oracle):
    """2048 bits, 12 000 hypergeometric instances (more than one wave of four-lane cooperative CTAs): the launcher
    picks the TWO-lane cooperative Frobenius kernel (CTAs of 120 instances, ragged last CTA, work units of 8
    coefficients); sampled instances at CTA / warp edges against the oracle, both exponents."""
    import torch

    import bench
    from cascade_b200 import _lib, api
    from cascade_b200.format import FLAG_NAN, Balls, from_complex, instance_major_to_entry_major, pack_entries

    lib = _lib.lib()
    L, batch, n = 64, 12000, 20
    dev = torch.device("cuda:0")
    drv = api.HostBufferDriver(0)
    a, b, c = bench.hypgeom_params(batch, L, seed=0x5EED0064)
    ode = api.acb_ode_hypgeom_batch(a, b, c, driver=drv)
    rho1 = api.ew(api.OP_SUB, from_complex([1] * batch, L), c, driver=drv)
    rho0 = Balls.zeros(2 * batch, L)
    om, ot = pack_entries(instance_major_to_entry_major(ode, batch, 9), 9, 2 * batch)
    d_om, d_ot = torch.from_numpy(om.view(np.int32)).to(dev), torch.from_numpy(ot).to(dev)
    d_res_m = torch.empty((n + 1, L, 2 * batch), dtype=torch.int32, device=dev)
    d_res_t = torch.empty((n + 1, 2 * batch, 4), dtype=torch.int32, device=dev)
    wsb = lib.cb200_frobenius_workspace_bytes(L, batch)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    sample = _sample(batch, [1, 15, 16, 119, 120, 121, 239, 240, 5999, 6000, 11879, 11880, 11998], 64)
    sub_ode = Balls.concat([ode[18 * i:18 * i + 18] for i in sample])
    for name, rho in (("rho = 0", rho0), ("rho = 1 - c", rho1)):
        rm, rt = pack_entries(rho, 1, 2 * batch)
        d_rm, d_rt = torch.from_numpy(rm.view(np.int32)).to(dev), torch.from_numpy(rt).to(dev)
        rc = lib.cb200_frobenius1_batch_dev(L, 2, 2, -1, n, batch, p(d_om), p(d_ot), 0, p(d_rm), p(d_rt), 0, p(d_res_m), p(d_res_t),
                                            p(ws), wsb, None)
        assert rc == 0
        torch.cuda.synchronize()
        assert int((d_res_t[:, :, 1] & FLAG_NAN).sum().item()) == 0, name
        got = _pick(d_res_m, d_res_t, sample, L, n + 1)
        sub_rho = Balls.concat([rho[2 * i:2 * i + 2] for i in sample])
        want = oracle.frobenius1_batch(2, 2, n, sub_ode, sub_rho, len(sample), False, nthreads=8)
        assert got.same_as(want), f"2048-bit Frobenius sweep, {name}: GPU run differs from the oracle on the sampled instances"


def test_config2_frobenius_sweep_full_size(oracle):
    """BASELINE configs[2]: Gauss hypergeometric acb_ode_solve_frobenius (M = 1) at 4096 bits, 16 384 (a,b,c) triples,
    256 terms, both exponents (rho = 0 and rho = 1 - c) -- reference src/frobenius_solver.c:72-121."""
    import torch

    import bench
    from cascade_b200 import _lib, api
    from cascade_b200.format import FLAG_NAN, Balls, from_complex, instance_major_to_entry_major, pack_entries

    lib = _lib.lib()
    L, batch, n = 128, 16384, 256
    dev = torch.device("cuda:0")
    drv = api.HostBufferDriver(0)
    a, b, c = bench.hypgeom_params(batch, L)
    ode = api.acb_ode_hypgeom_batch(a, b, c, driver=drv)
    rho1 = api.ew(api.OP_SUB, from_complex([1] * batch, L), c, driver=drv)
    rho0 = Balls.zeros(2 * batch, L)
    om, ot = pack_entries(instance_major_to_entry_major(ode, batch, 9), 9, 2 * batch)
    d_om, d_ot = torch.from_numpy(om.view(np.int32)).to(dev), torch.from_numpy(ot).to(dev)
    d_res_m = torch.empty((n + 1, L, 2 * batch), dtype=torch.int32, device=dev)
    d_res_t = torch.empty((n + 1, 2 * batch, 4), dtype=torch.int32, device=dev)
    wsb = lib.cb200_frobenius_workspace_bytes(L, batch)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    # block edges of every block size the launcher may pick (112 all-shared, 192 hybrid, 64-instance cooperative units)
    sample = _sample(batch, [1, 31, 32, 63, 64, 111, 112, 113, 191, 192, 255, 256, 8191, 8192, 16271, 16272], 21)
    sub_ode = Balls.concat([ode[18 * i:18 * i + 18] for i in sample])
    for name, rho in (("rho = 0", rho0), ("rho = 1 - c", rho1)):
        rm, rt = pack_entries(rho, 1, 2 * batch)
        d_rm, d_rt = torch.from_numpy(rm.view(np.int32)).to(dev), torch.from_numpy(rt).to(dev)
        rc = lib.cb200_frobenius1_batch_dev(L, 2, 2, -1, n, batch, p(d_om), p(d_ot), 0, p(d_rm), p(d_rt), 0, p(d_res_m), p(d_res_t),
                                            p(ws), wsb, None)
        assert rc == 0
        torch.cuda.synchronize()
        assert int((d_res_t[:, :, 1] & FLAG_NAN).sum().item()) == 0, name
        got = _pick(d_res_m, d_res_t, sample, L, n + 1)
        sub_rho = Balls.concat([rho[2 * i:2 * i + 2] for i in sample])
        want = oracle.frobenius1_batch(2, 2, n, sub_ode, sub_rho, len(sample), False, nthreads=8)
        assert got.same_as(want), f"configs[2] {name}: full-size GPU run differs from the oracle on the sampled instances"
        if name == "rho = 0":
            # the reference's acceptance test (tests/singleton_frobenius.c: residual contains zero), rho = 0: no shift
            for k in (0, len(sample) - 1):
                series = got[k * (n + 1) * 2:(k + 1) * (n + 1) * 2]
                assert oracle.ode_solves(2, 2, sub_ode[18 * k:18 * k + 18], series, n - 2)


def test_config3_legendre_tile_full_size(oracle):
    """BASELINE configs[3]: one 2^17-instance tile of the Legendre sweep, 1000 terms at 1024 bits, through the
    integer-operator entry point (reference src/fuchs_solver.c:96-145 on src/examples.c:3-11 operators); the launcher
    picks the persistent form by itself at this size.  The tile starts at n = 917 504: the LAST tile of the sweep
    n = 0 .. 2^20 - 1, whose multipliers n(n+1) are the largest."""
    import torch

    from cascade_b200 import _lib, api
    from cascade_b200.format import FLAG_NAN, from_complex

    lib = _lib.lib()
    L, batch, n, n0 = 32, 1 << 17, 1000, (1 << 20) - (1 << 17)
    dev = torch.device("cuda:0")
    ns = torch.arange(n0, n0 + batch, dtype=torch.int64, device=dev)
    oi = torch.zeros((9, batch), dtype=torch.int64, device=dev)
    oi[0] = ns * (ns + 1)
    oi[4] = -2
    oi[6] = 1
    oi[8] = -1
    d_res_m = torch.zeros((n + 1, L, 2 * batch), dtype=torch.int32, device=dev)
    d_res_t = torch.zeros((n + 1, 2 * batch, 4), dtype=torch.int32, device=dev)
    d_res_t[:, :, 1] = 2
    d_res_m[0:2, L - 1, 0::2] = -(2 ** 31)  # c0 = c1 = 1
    d_res_t[0:2, 0::2, 0] = 1
    d_res_t[0:2, 0::2, 1] = 0
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.cb200_fuchs_intop_batch_dev(L, 2, 2, -2, n, batch, p(oi), 0, p(d_res_m), p(d_res_t), None)
    assert rc == 0
    torch.cuda.synchronize()
    assert int((d_res_t[:, :, 1] & FLAG_NAN).sum().item()) == 0
    assert int((d_res_t[:, 1::2, 2] != 0).sum().item()) == 0          # a real sweep: imaginary parts stay exact zeros
    sample = _sample(batch, [1, 31, 32, 127, 128, 255, 256, 383, 384, 385, 511, 512, 56831, 56832, 131071 - 384], 22)
    got = _pick(d_res_m, d_res_t, sample, L, n + 1)
    ode = api.acb_ode_legendre_batch([n0 + i for i in sample], L)
    init = from_complex([1, 1] * len(sample), L)
    rc, want = oracle.fuchs_batch(2, 2, n, ode, init, len(sample), False, nthreads=8)
    assert rc == 0
    assert got.same_as(want), "configs[3]: full-size tile differs from the oracle (general Fuchs loop) on the sampled instances"
    for k in (0, len(sample) - 1):
        assert oracle.ode_solves(2, 2, ode[18 * k:18 * k + 18], got[k * (n + 1) * 2:(k + 1) * (n + 1) * 2], 300)


def test_config4_horner_full_size(oracle):
    """BASELINE configs[4], evaluation leg: one 3 504-term series at 2^18 points, 2048 bits (acb_poly_evaluate inside
    acb_ode_solution_evaluate, reference src/acb_ode_solution.c:40); the launcher picks the kernel form."""
    import torch

    from cascade_b200 import _lib
    from cascade_b200.format import FLAG_NAN, Balls, pack_entries, random_balls

    lib = _lib.lib()
    L, length, npts = 64, 3504, 1 << 18
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0x5EED0005)
    f = random_balls(rng, 2 * length, L, -1, 1, "ulp")
    pts = random_balls(rng, 2 * npts, L, -3, -2, "zero")
    fm, ft = pack_entries(f, length, 2)
    pm, pt = pack_entries(pts, 1, 2 * npts)
    d = lambda x: torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).to(dev)
    d_fm, d_ft, d_pm, d_pt = d(fm), d(ft), d(pm), d(pt)
    d_om = torch.empty((1, L, 2 * npts), dtype=torch.int32, device=dev)
    d_ot = torch.empty((1, 2 * npts, 4), dtype=torch.int32, device=dev)
    wsb = lib.cb200_horner_workspace_bytes(L, npts)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.cb200_horner_batch_dev(L, length, npts, 1, p(d_fm), p(d_ft), 1, p(d_pm), p(d_pt), p(d_om), p(d_ot), p(ws), wsb, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert int((d_ot[:, :, 1] & FLAG_NAN).sum().item()) == 0
    sample = _sample(npts, [1, 31, 32, 63, 64, 191, 192, 255, 256, 257, 37887, 37888, npts - 256], 23)
    got = _pick(d_om, d_ot, sample, L, 1)
    sub = Balls.concat([pts[2 * i:2 * i + 2] for i in sample])
    want = oracle.horner_batch(f, sub, 1, nthreads=8)
    assert got.same_as(want), "configs[4] Horner: full-size GPU run differs from the oracle on the sampled points"


def test_config4_continuation_full_size(oracle):
    """BASELINE configs[4], continuation leg as a device-resident batch: 16 384 solutions of the hypergeometric operator,
    two path steps of 3 504 terms each at 2048 bits (analytic_continuation, reference src/fuchs_solver.c:147-163: shift,
    acb_ode_solve_fuchs, re-expansion); the launcher picks the recurrence / Horner kernel forms."""
    import torch

    from fractions import Fraction as F

    from cascade_b200 import _lib, api
    from cascade_b200.format import FLAG_NAN, Balls, from_complex, instance_major_to_entry_major, pack_entries, random_balls

    lib = _lib.lib()
    L, n, batch = 64, 3504, 16384
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0x5EED0006)
    par = random_balls(rng, 6, L, -1, 1, "zero")
    ode = oracle.example("hypgeom", L, 0, par)
    path = from_complex([(F(1, 2), F(1, 8)), (F(9, 16), F(3, 16)), (F(5, 8), F(1, 8))], L)
    y0 = random_balls(rng, batch * 4, L, -1, 0, "zero")
    om, ot = api._ode_to_device(ode, 2, 2, batch, True)
    pm, pt = pack_entries(path, 3, 2)
    ym, yt = pack_entries(instance_major_to_entry_major(y0, batch, 2), 2, 2 * batch)
    d = lambda x: torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).to(dev)
    d_om, d_ot, d_pm, d_pt, d_ym, d_yt = d(om), d(ot), d(pm), d(pt), d(ym), d(yt)
    wsb = lib.cb200_continuation_workspace_bytes(L, 2, 2, n, batch)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.cb200_continuation_dev(L, 2, 2, n, batch, 3, p(d_om), p(d_ot), p(d_pm), p(d_pt), None, p(d_ym), p(d_yt), p(ws), wsb, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert int((d_yt[:, :, 1] & FLAG_NAN).sum().item()) == 0
    sample = _sample(batch, [1, 31, 32, 95, 96, 97, 191, 192, 255, 256, 14207, 14208], 24, extra=2)
    got = _pick(d_ym, d_yt, sample, L, 2)
    sub = Balls.concat([y0[4 * i:4 * i + 4] for i in sample])
    want = oracle.continuation(2, 2, n, ode, path, sub, len(sample))
    assert got.same_as(want), "configs[4] continuation: full-size GPU run differs from the oracle on the sampled solutions"
