"""Device-resident timings of the other BASELINE.json configs (development report, not the bench
contract): config 3 (hypergeometric Frobenius sweep, 4096 bits), config 4 (Legendre sweep, 1024
bits, per-instance integer operators), config 5 (Horner evaluation at many points, 2048 bits).
Random dense inputs of the right shape; sizes can be scaled down with --scale."""
import argparse
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from cascade_b200 import _lib, api
from cascade_b200.format import instance_major_to_entry_major, pack_entries, random_balls

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--only", type=int, default=0, help="run only this config (2 = per-instance-operator sweep, 3, 4 or 5)")
ap.add_argument("--lean", action="store_true", help="only the main kernel of each config (for ncu)")
ap.add_argument("--terms", type=int, default=0, help="override the number of terms (short runs for ncu)")
ap.add_argument("--continuation", action="store_true", help="also time the (slow, sequential) continuation leg")
ap.add_argument("--continuation-batch", type=int, default=2, help="initial vectors continued along the path (2 = one fundamental system)")
args = ap.parse_args()
lib = _lib.lib()
dev = torch.device("cuda:0")
p = lambda t: C.c_void_p(t.data_ptr())


def dev_arr(mant, meta):
    return torch.from_numpy(mant.view(np.int32)).to(dev), torch.from_numpy(meta).to(dev)


def rand_dev(entries, L, S, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    m = torch.randint(0, 2**31 - 1, (entries, L, S), device=dev, dtype=torch.int32, generator=g) * 2 + 1
    m[:, L - 1, :] |= -(2**31)
    t = torch.zeros((entries, S, 4), device=dev, dtype=torch.int32)
    return m, t


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn()
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0, rc
        best = min(best, e0.elapsed_time(e1))
    return best


class _Out(list):
    def append(self, o):
        print(json.dumps(o), flush=True)
        super().append(o)


out = _Out()
def config3():
    # ---- config 3: Frobenius, hypergeometric sweep, 4096 bits, per-instance (a, b, c)
    L, n = 128, (args.terms or 256)
    batch = max(64, int(16384 * args.scale))
    rng = np.random.default_rng(3)
    small = 64
    par = [random_balls(rng, small * 2, L, -1, 1, "zero") for _ in range(3)]
    odes = api.acb_ode_hypgeom_batch(*par)
    om, ot = pack_entries(instance_major_to_entry_major(odes, small, 9), 9, 2 * small)
    reps = batch // small
    d_om = torch.from_numpy(np.tile(om.view(np.int32), (1, 1, reps))).to(dev)
    d_ot = torch.from_numpy(np.tile(ot, (1, reps, 1))).to(dev)
    d_rm, d_rt = torch.zeros((1, L, 2 * batch), dtype=torch.int32, device=dev), torch.zeros((1, 2 * batch, 4), dtype=torch.int32, device=dev)
    d_rt[:, :, 1] = 2  # rho = 0
    d_res_m = torch.empty((n + 1, L, 2 * batch), dtype=torch.int32, device=dev)
    d_res_t = torch.empty((n + 1, 2 * batch, 4), dtype=torch.int32, device=dev)
    wsb = lib.cb200_frobenius_workspace_bytes(L, batch)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ms = timed(lambda: lib.cb200_frobenius1_batch_dev(L, 2, 2, -1, n, batch, p(d_om), p(d_ot), 0, p(d_rm), p(d_rt), 0,
                                                      p(d_res_m), p(d_res_t), p(ws), wsb, None))
    coeffs = batch * n
    out.append({"config": 3, "what": f"hypgeom Frobenius M=1, rho = 0, 4096 bits, {batch} (a,b,c) triples, {n} terms", "ms": ms,
                "ball_coeffs_per_s": coeffs / ms * 1e3, "T_limb_MAC_per_s_36L2": coeffs * 36 * L * L / ms * 1e3 / 1e12})
    # second exponent of the sweep: rho = 1 - c is a dense complex ball per instance
    d_r2m, d_r2t = rand_dev(1, L, 2 * batch, 31)
    ms = timed(lambda: lib.cb200_frobenius1_batch_dev(L, 2, 2, -1, n, batch, p(d_om), p(d_ot), 0, p(d_r2m), p(d_r2t), 0,
                                                      p(d_res_m), p(d_res_t), p(ws), wsb, None))
    out.append({"config": 3, "what": f"hypgeom Frobenius M=1, dense complex rho per instance (the rho = 1 - c run), 4096 bits, {batch} triples, {n} terms",
                "ms": ms, "ball_coeffs_per_s": coeffs / ms * 1e3, "T_limb_MAC_per_s_36L2": coeffs * 36 * L * L / ms * 1e3 / 1e12})
    del d_res_m, d_res_t, ws
    torch.cuda.empty_cache()


def config4():
    # ---- config 4: Legendre sweep, per-instance integer operators, real data (integer-operator kernel)
    L, n = 32, (args.terms or 1000)
    # the full 2^20 x 1000 sweep is 302 GB of complex results: more than one B200 holds; cap at 2^17 instances per launch
    batch = max(1024, min(2**17, int(2**20 * args.scale)))
    ns = np.arange(batch, dtype=np.int64)
    ode_int = np.zeros((9, batch), np.int64)
    ode_int[0] = ns * (ns + 1)
    ode_int[4] = -2
    ode_int[6] = 1
    ode_int[8] = -1
    d_oi = torch.from_numpy(ode_int).to(dev)
    d_res_m = torch.zeros((n + 1, L, 2 * batch), dtype=torch.int32, device=dev)
    d_res_t = torch.zeros((n + 1, 2 * batch, 4), dtype=torch.int32, device=dev)
    d_res_t[:, :, 1] = 2
    d_res_m[0:2, L - 1, 0::2] = -(2**31)  # c0 = c1 = 1
    d_res_t[0:2, 0::2, 0] = 1
    d_res_t[0:2, 0::2, 1] = 0
    ms = timed(lambda: lib.cb200_fuchs_intop_batch_dev(L, 2, 2, -2, n, batch, p(d_oi), 0, p(d_res_m), p(d_res_t), None), reps=2)
    coeffs = batch * (n - 1)
    out.append({"config": 4, "what": f"Legendre sweep n=0..{batch - 1}, 1024 bits, {n} terms, integer-operator kernel",
                "ms": ms, "ball_coeffs_per_s": coeffs / ms * 1e3, "GB_per_s_written": coeffs * 288 / ms * 1e3 / 1e9})
    del d_res_m, d_res_t
    torch.cuda.empty_cache()


def config5():
    # ---- config 5: Horner at many points, 2048 bits, one 3504-term series
    L, length = 64, (args.terms or 3504)
    npts = max(1024, int(10**6 * args.scale))
    f_m, f_t = rand_dev(length, L, 2, 5)
    pts_m, pts_t = rand_dev(1, L, 2 * npts, 6)
    pts_t[:, :, 0] = -2  # |x| < 1/4
    o_m = torch.empty((1, L, 2 * npts), dtype=torch.int32, device=dev)
    o_t = torch.empty((1, 2 * npts, 4), dtype=torch.int32, device=dev)
    wsb = lib.cb200_horner_workspace_bytes(L, npts)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ms = timed(lambda: lib.cb200_horner_batch_dev(L, length, npts, 1, p(f_m), p(f_t), 1, p(pts_m), p(pts_t), p(o_m), p(o_t), p(ws), wsb, None), reps=1)
    steps = npts * (length - 1)
    out.append({"config": 5, "what": f"Horner evaluation of a {length}-term series at {npts} points, 2048 bits", "ms": ms,
                "horner_steps_per_s": steps / ms * 1e3, "T_limb_MAC_per_s_4L2": steps * 4 * L * L / ms * 1e3 / 1e12})

def config2_sweep():
    # ---- configs[1] with an operator per instance (parameter sweep): per-instance table + recurrence
    L, n = 32, (args.terms or 99)
    batch = max(1024, int(66304 * args.scale))
    om, ot = rand_dev(9, L, 2 * batch, 21)
    rm, rt = rand_dev(n + 1, L, 2 * batch, 22)
    wsb = lib.cb200_fuchs_workspace_bytes(L, 2, 2, -2, n, batch, 0)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ms = timed(lambda: lib.cb200_fuchs_batch_dev(L, 2, 2, -2, n, batch, p(om), p(ot), 0, p(rm), p(rt), p(ws), wsb, None))
    coeffs = batch * (n - 1)
    out.append({"config": "2-sweep", "what": f"configs[1] with an operator per instance, 1024 bits, {batch} instances, {n} terms",
                "ms": ms, "ball_coeffs_per_s": coeffs / ms * 1e3, "T_limb_MAC_per_s_24L2": coeffs * 24 * L * L / ms * 1e3 / 1e12})
    del ws
    torch.cuda.empty_cache()


def config3_log():
    # ---- config 3, logarithmic variant: acb_ode_solve_frobenius with M = 2 (polynomial-in-rho recurrence), 1024 bits
    L, n, M = 32, 48, 2
    batch = max(256, int(16384 * args.scale))
    rng = np.random.default_rng(33)
    small = 64
    par = [random_balls(rng, small * 2, L, -1, 1, "zero") for _ in range(3)]
    odes = api.acb_ode_hypgeom_batch(*par)
    om, ot = pack_entries(instance_major_to_entry_major(odes, small, 9), 9, 2 * small)
    reps = batch // small
    d_om = torch.from_numpy(np.tile(om.view(np.int32), (1, 1, reps))).to(dev)
    d_ot = torch.from_numpy(np.tile(ot, (1, reps, 1))).to(dev)
    d_rm, d_rt = rand_dev(1, L, 2 * batch, 34)  # generic complex rho per instance
    d_gm = torch.empty((M * (n + 1), L, 2 * batch), dtype=torch.int32, device=dev)
    d_gt = torch.empty((M * (n + 1), 2 * batch, 4), dtype=torch.int32, device=dev)
    wsb = lib.cb200_frobenius_general_workspace_bytes(L, 2, 2, M, n, batch)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ms = timed(lambda: lib.cb200_frobenius_general_batch_dev(L, 2, 2, -1, n, batch, p(d_om), p(d_ot), 0, p(d_rm), p(d_rt), 0,
                                                             M, M, p(d_gm), p(d_gt), p(ws), wsb, None), reps=1)
    out.append({"config": "3-log", "what": f"hypgeom Frobenius M=2 (polynomials in rho), 1024 bits, {batch} triples, {n} terms",
                "ms": ms, "ball_coeffs_per_s": batch * n * M / ms * 1e3, "workspace_GB": wsb / 1e9})
    del d_gm, d_gt, ws
    torch.cuda.empty_cache()


def config5_extend():
    # ---- config 5, "acb_ode_solution_extend" leg: a batch of _acb_ode_solution_extend calls (Horner evaluation of a
    # polynomial in rho and of its derivatives: reference src/acb_ode_solution.c:97-109; polynomial length 20 as in
    # reference tests/solution_extend.c), 2048 bits, M = 3
    L, M, cap, fcap = 64, 3, 8, 20
    batch = max(1024, int(262144 * args.scale))
    g_m = torch.zeros((M * cap, L, 2 * batch), dtype=torch.int32, device=dev)
    g_t = torch.zeros((M * cap, 2 * batch, 4), dtype=torch.int32, device=dev)
    g_t[:, :, 1] = 2
    f_m, f_t = rand_dev(fcap, L, 2 * batch, 11)
    r_m, r_t = rand_dev(1, L, 2 * batch, 12)
    wsb = lib.cb200_solution_op_workspace_bytes(L, M, batch)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    f_keep_m, f_keep_t = f_m.clone(), f_t.clone()

    def run():
        f_m.copy_(f_keep_m)  # the call consumes its polynomial
        f_t.copy_(f_keep_t)
        return lib.cb200_solution_op_dev(0, L, batch, M, M, cap, 3, p(r_m), p(r_t), p(g_m), p(g_t), p(f_m), p(f_t), fcap,
                                         p(ws), wsb, None)
    ms = timed(run, reps=1)
    out.append({"config": "5-extend", "what": f"_acb_ode_solution_extend, M={M}, polynomial of {fcap} terms in rho, {batch} solutions, 2048 bits (includes restoring the consumed polynomial)",
                "ms": ms, "extends_per_s": batch / ms * 1e3})


def config5_eval():
    # ---- config 5, full acb_ode_solution_evaluate (Horner of M = 2 series + acb_log + acb_pow), 1024 bits
    L, cap, M = 32, 256, 2
    npts = max(1024, int(131072 * args.scale))
    g_m, g_t = rand_dev(M * cap, L, 2, 7)
    r_m, r_t = rand_dev(1, L, 2, 8)
    pts_m, pts_t = rand_dev(1, L, 2 * npts, 9)
    pts_t[:, :, 0] = -2
    o_m = torch.empty((1, L, 2 * npts), dtype=torch.int32, device=dev)
    o_t = torch.empty((1, 2 * npts, 4), dtype=torch.int32, device=dev)
    wsb = lib.cb200_evaluate_workspace_bytes(L, npts)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ms = timed(lambda: lib.cb200_solution_evaluate_dev(L, M, cap, npts, p(g_m), p(g_t), p(r_m), p(r_t), 2, 0, p(pts_m), p(pts_t),
                                                       p(o_m), p(o_t), p(ws), wsb, None), reps=1)
    out.append({"config": "5-eval", "what": f"acb_ode_solution_evaluate, M=2, {cap}-term series, generic rho (log + pow), {npts} points, 1024 bits",
                "ms": ms, "points_per_s": npts / ms * 1e3})


if args.only in (0, 2):
    config2_sweep()
if args.only in (0, 3):
    config3()
    if not args.lean:
        config3_log()
if args.only in (0, 5) and not args.lean:
    config5_eval()
    config5_extend()
if args.only in (0, 4):
    config4()
if args.only in (0, 5):
    config5()
if args.continuation:
    # ---- config 5, continuation leg (slow: sequential; run with --continuation): 64-corner path, ~3504 terms per step, 2048 bits, `order` solutions
    L, n, steps = 64, 3504, max(4, int(64 * min(1.0, args.scale * 4)))
    rng = np.random.default_rng(5)
    par = random_balls(rng, 6, L, -1, 1, "zero")
    ode = api.acb_ode_hypgeom_batch(par[0:2], par[2:4], par[4:6])
    from fractions import Fraction as F
    from cascade_b200.format import from_complex
    th = np.linspace(0.0, 2 * np.pi, steps + 1)
    path = from_complex([(F(float(0.5 + 0.25 * np.cos(t))).limit_denominator(2**30), F(float(0.25 * np.sin(t))).limit_denominator(2**30)) for t in th], L)
    B = args.continuation_batch
    if B == 2:
        y0 = from_complex([1, 0, 0, 1], L)
    else:
        y0 = random_balls(rng, 4 * B, L, -1, 0, "zero")
    om, ot = api._ode_to_device(ode, 2, 2, 2, True)
    pm, pt = pack_entries(path, steps + 1, 2)
    ym, yt = pack_entries(instance_major_to_entry_major(y0, B, 2), 2, 2 * B)
    d_om, d_ot = dev_arr(om, ot)
    d_pm, d_pt = dev_arr(pm, pt)
    d_ym, d_yt = dev_arr(ym, yt)
    wsb = lib.cb200_continuation_workspace_bytes(L, 2, 2, n, B)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ms = timed(lambda: lib.cb200_continuation_dev(L, 2, 2, n, B, steps + 1, p(d_om), p(d_ot), p(d_pm), p(d_pt), None, p(d_ym), p(d_yt), p(ws), wsb, None), reps=1)
    out.append({"config": 5, "what": f"analytic continuation, {steps}-step path, {n} terms per step, 2048 bits, {B} solutions"
                                     + (" (sequential: latency bound)" if B == 2 else " continued together (device-resident pipeline)"),
                "ms": ms, "ball_coeffs_per_s": B * steps * (n - 1) / ms * 1e3, "ms_per_step": ms / steps, "workspace_GB": wsb / 1e9})
