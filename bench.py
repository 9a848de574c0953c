#!/usr/bin/env python
"""bench.py -- throughput of the batched Cascade hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config {1,2,3,4}] [--impl reference]
    (N > 1: launched by torchrun, one rank per GPU)

--config selects one of BASELINE.json's configs (0-based index as in the file; default 1 = the headline):

  1  Bessel ODE (complex nu) shifted to a complex ordinary point, 2000-term Fuchs series at 1024 bits, 65536
     initial-condition pairs per GPU, one operator shared by the batch (acb_ode_solve_fuchs).      roofline: IMAD, 24 L^2
  2  Gauss hypergeometric acb_ode_solve_frobenius (M = 1) at 4096 bits, 16384 (a,b,c) triples per GPU, 256 terms,
     exponents rho = 0 and rho = 1 - c (one step = both runs).                                     roofline: IMAD, 36 L^2
  3  Legendre sweep n = 0 .. 2^20-1, 1000 terms at 1024 bits (integer-operator kernel), the WHOLE sweep per step, tiled
     (2^17 instances per launch; the 302 GB of series do not fit one GPU and stay device-side per tile), sharded over
     the ranks, with an NCCL all-gather of the last coefficient of every instance inside the step.  roofline: HBM
  4  Horner evaluation of one 3504-term series (the length truncation_order gives config 4's continuation steps) at
     10^6 points per GPU, 2048 bits (acb_poly_evaluate inside acb_ode_solution_evaluate / the Taylor re-expansion
     of analytic_continuation).                                                                    roofline: IMAD, 4 L^2

A "step" is one full pass of that workload.  Metric: ball-coefficients per second (complex series coefficients
written -- configs 1-3 -- or consumed by Horner steps -- config 4 -- / time).

  value : device-resident (inputs already in HBM, CUDA events around the kernels)
  e2e   : the same metric through the host-buffer C-ABI call (cb200_*_host): host inputs copied H2D and the results
          copied D2H inside the timed region
  roofline : the bound of the dominant kernel.  IMAD: achieved = SURVEY 8(d)'s schoolbook limb multiply-adds per
          coefficient x coefficients / kernel time, peak = measured plain IMAD.WIDE rate (profiles/imad_peak_r01.json);
          HBM: algorithmic bytes written / kernel time, peak = MEASURED_PEAKS.json.  traffic = DRAM bytes per launch
          from the committed ncu capture of the same config (profiles/traffic_r02.json), else null.
  cpu_baseline : the oracle port (Cascade's loops on GMP mpn) on the host cores, bounded sample.

--impl reference times that CPU port alone (Cascade + Arb itself cannot be built in this image:
no Arb/FLINT/GMP headers), same metric/config, JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- BASELINE configs[1] (kept as module constants: tests/test_gpu_fullsize.py uses them)
L = 32            # 1024 bits
ORDER = DEGREE = 2
N_TERMS = 2000    # coefficients 0..2000; 1999 new ones per instance
BATCH = 65536     # per GPU
SEED = 0x5EED0002
UNIT = "ball-coeffs/s"
METRIC = "ball-coeffs/sec, 1024-bit batched acb_ode_solve_fuchs"


def make_inputs(batch: int, n: int, drv_for_setup=None):
    """configs[1]: Bessel(nu) shifted to a, and `batch` random initial-value pairs (host Balls)."""
    from cascade_b200 import api
    from cascade_b200.format import random_balls
    rng = np.random.default_rng(SEED)
    nu = random_balls(rng, 2, L, 0, 0, "zero")
    a = random_balls(rng, 2, L, 0, 1, "zero")
    ode0, _ = api.acb_ode_bessel_batch(nu, driver=drv_for_setup)
    ode = api.acb_ode_shift_batch(ode0, a, ORDER, DEGREE, driver=drv_for_setup)
    init = random_balls(rng, batch * 2 * 2, L, -1, 0, "zero")
    return ode, init


def hypgeom_params(batch: int, Lh: int, seed: int = 0x5EED0003):
    """configs[2]: (a, b, c) complex with unit-square midpoints and full mantissas, radius 0; c is kept away from the
    integers (|Im c| >= 2^-8) so that rho = 0 and rho = 1 - c are simple exponents."""
    from cascade_b200.format import random_balls
    rng = np.random.default_rng(seed)
    a, b, c = (random_balls(rng, 2 * batch, Lh, -1, 0, "zero") for _ in range(3))
    c.meta[1::2, 0] = np.maximum(c.meta[1::2, 0], -7)
    return a, b, c


def clocks_sampler(stop, out, gpu_index):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def summarise_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_max_mhz": int(samples[0][1]) if samples[0][1].isdigit() else None, "reasons": reasons}


def pinned_array(lib, shape, dtype):
    """A page-locked numpy array of exactly this size (cb200_host_alloc); falls back to pageable memory."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.cb200_host_alloc(n)
    if not p:
        return np.zeros(shape, dtype), False
    buf = (C.c_uint8 * n).from_address(p)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), True


# ====================================================================== CPU port legs (oracle; never on the product path)
def cpu_port(cfg: int, ncores: int, target_seconds: float, args):
    """The oracle port of the same workload on `ncores` host threads, bounded sample.  Returns (coeffs/s, sample text)."""
    from oracle import cbo
    from cascade_b200.format import Balls, from_complex, random_balls

    def sized(run_one, per_unit_coeffs, unit_name):
        dt1 = run_one(ncores)                      # one unit per core
        rounds = max(1, min(64, int(target_seconds / max(dt1, 1e-3))))
        units = ncores * rounds
        dt = run_one(units)
        return units * per_unit_coeffs / dt, f"{units} {unit_name}, {dt:.1f} s on {ncores} thread(s)"

    if cfg == 1:
        rng = np.random.default_rng(SEED)
        nu = random_balls(rng, 2, L, 0, 0, "zero")
        a = random_balls(rng, 2, L, 0, 1, "zero")
        ode = cbo.ode_shift(ORDER, DEGREE, cbo.example("bessel", L, 0, nu), a)
        n = args.terms or N_TERMS

        def run(batch):
            init = random_balls(rng, batch * 4, L, -1, 0, "zero")
            res = cbo._series_with_init(init, batch, n, 2)
            t0 = time.perf_counter()
            rc = cbo.lib().cbo_fuchs_batch(L, ORDER, DEGREE, n, batch, 1, ode.mant, ode.meta, res.mant, res.meta, ncores)
            assert rc == 0
            return time.perf_counter() - t0
        v, s = sized(run, n - 1, f"instances x {n - 1} coefficients")
    elif cfg == 2:
        Lh, n = 128, args.terms or 256

        def run(batch):
            a, b, c = hypgeom_params(batch, Lh, seed=77)
            ode = Balls.concat([cbo.example("hypgeom", Lh, 0, Balls.concat([a[2 * i:2 * i + 2], b[2 * i:2 * i + 2], c[2 * i:2 * i + 2]]))
                                for i in range(batch)])
            rho = Balls.zeros(2 * batch, Lh)
            t0 = time.perf_counter()
            cbo.frobenius1_batch(2, 2, n, ode, rho, batch, False, nthreads=ncores)
            return time.perf_counter() - t0
        v, s = sized(run, n, f"(a,b,c) triples x {n} coefficients (rho = 0)")
    elif cfg == 3:
        n = args.terms or 1000

        def run(batch):
            ode = Balls.concat([cbo.example("legendre", L, 1000 + k) for k in range(batch)])
            init = from_complex([1, 1] * batch, L)
            t0 = time.perf_counter()
            rc, _ = cbo.fuchs_batch(2, 2, n, ode, init, batch, False, nthreads=ncores)
            assert rc == 0
            return time.perf_counter() - t0
        v, s = sized(run, n - 1, f"Legendre operators x {n - 1} coefficients")
    else:
        Lh, length = 64, args.terms or 3504
        rng = np.random.default_rng(5)
        f = random_balls(rng, 2 * length, Lh, -1, 1, "zero")

        def run(npts):
            pts = random_balls(rng, 2 * npts, Lh, -3, -2, "zero")
            t0 = time.perf_counter()
            cbo.horner_batch(f, pts, 1, nthreads=ncores)
            return time.perf_counter() - t0
        v, s = sized(run, length - 1, f"points x {length - 1} Horner steps")
    return v, s + ", oracle port (Cascade loops on GMP 6.3 mpn); Cascade+Arb itself is not buildable here"


# ====================================================================== workloads (device side)
class Workload:
    """One BASELINE config: resident inputs, the device step, the host-buffer (e2e) step, and its roofline."""
    scaling = "weak"

    def __init__(self, lib, api, torch, dev, rank, world, args):
        self.lib, self.api, self.torch, self.dev, self.rank, self.world, self.args = lib, api, torch, dev, rank, world, args
        self.stream = torch.cuda.current_stream(dev)
        self.sp = C.c_void_p(self.stream.cuda_stream)

    def p(self, t):
        return C.c_void_p(t.data_ptr())

    def to_dev(self, m, t):
        return (self.torch.from_numpy(m.view(np.int32)).to(self.dev), self.torch.from_numpy(t).to(self.dev))


class FuchsShared(Workload):
    cfg = 1
    dtype = "u32 limbs (1024-bit ball)"

    def setup(self):
        from cascade_b200.format import instance_major_to_entry_major, pack_entries, random_balls
        torch, lib, api, dev, args = self.torch, self.lib, self.api, self.dev, self.args
        self.batch, self.n = args.batch or BATCH, args.terms or N_TERMS
        batch, n = self.batch, self.n
        S = 2 * batch
        ode, init = make_inputs(batch, n, api.HostBufferDriver(dev.index))
        if self.rank:  # different initial values on every rank (independent instances; weak scaling)
            init = random_balls(np.random.default_rng(SEED + self.rank), batch * 4, L, -1, 0, "zero")
        self.ode_m, self.ode_t = api._ode_to_device(ode, ORDER, DEGREE, batch, True)
        self.init_m, self.init_t = pack_entries(instance_major_to_entry_major(init, batch, 2), 2, S)
        self.d_ode_m, self.d_ode_t = self.to_dev(self.ode_m, self.ode_t)
        self.d_init_m, self.d_init_t = self.to_dev(self.init_m, self.init_t)
        self.d_res_m = torch.empty((n + 1, L, S), dtype=torch.int32, device=dev)
        self.d_res_t = torch.empty((n + 1, S, 4), dtype=torch.int32, device=dev)
        self.ws_bytes = lib.cb200_fuchs_workspace_bytes(L, 2, 2, -2, n, batch, 1)
        self.d_ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.coeffs = batch * (n - 1)
        self.config = self.describe()

    metric = METRIC

    def describe(self):
        batch, n = self.args.batch or BATCH, self.args.terms or N_TERMS
        return {"workload": "configs[1]: Bessel ODE (complex nu) shifted to an ordinary point, "
                            f"{n}-term Fuchs series, 1024 bits, {batch} initial conditions per GPU, shared operator",
                "L_limbs": L, "order": ORDER, "degree": DEGREE, "terms": n, "batch_per_gpu": batch, "seed": hex(SEED),
                "cache": "outputs (tens of GB per step) and the coefficient window far exceed the 126 MB L2; no flush needed"}

    def prepare(self):
        self.d_res_m[:2].copy_(self.d_init_m)
        self.d_res_t[:2].copy_(self.d_init_t)

    def kernels(self):
        p = self.p
        rc = self.lib.cb200_fuchs_batch_dev(L, ORDER, DEGREE, -2, self.n, self.batch, p(self.d_ode_m), p(self.d_ode_t), 1,
                                            p(self.d_res_m), p(self.d_res_t), p(self.d_ws), self.ws_bytes, self.sp)
        assert rc == 0, f"cb200_fuchs_batch_dev -> {rc}"

    def finish(self):
        pass

    def roofline(self, kernel_ms, peaks, imad_peak):
        macs = 24 * L * L * self.coeffs  # (4W+8) L^2 per coefficient, W = 4
        return imad_roofline(macs, kernel_ms, imad_peak, self.coeffs * 2 * (4 * L + 16), peaks,
                             traffic_for(1, {"batch": self.batch, "terms": self.n}))

    def e2e_host_bytes(self):
        return (self.n + 1) * 2 * self.batch * (4 * L + 16)

    def e2e_setup(self, tiles=1):
        """Host side of the end-to-end leg: the caller's FULL result arrays, page-locked.  When the host cannot hold
        them for every rank of the box (`tiles` > 1, chosen by main from the free memory) the batch goes through the
        same host call in `tiles` instance tiles whose result arrays are reused -- the same coefficients, the same bytes
        over PCIe, a caller that consumes one tile before asking for the next."""
        from cascade_b200.format import pack_entries
        lib, n = self.lib, self.n
        self.e2e_tiles = tiles
        tb = self.batch // tiles
        S = 2 * tb
        self.h_res_m, pm = pinned_array(lib, (n + 1, L, S), np.uint32)
        self.h_res_t, pt = pinned_array(lib, (n + 1, S, 4), np.int32)
        # per-tile initial rows (slots of a tile are contiguous in the [entry][limb][slot] layout of the whole batch)
        self.h_init = [(np.ascontiguousarray(self.init_m[:, :, 2 * tb * k:2 * tb * (k + 1)]),
                        np.ascontiguousarray(self.init_t[:, 2 * tb * k:2 * tb * (k + 1)])) for k in range(tiles)]
        # the device-resident leg's buffers are not needed any more: the host call owns its own (cached) ones
        del self.d_res_m, self.d_res_t, self.d_ws
        self.torch.cuda.empty_cache()
        self.h2d = tiles * (self.ode_m.nbytes + self.ode_t.nbytes) + 2 * L * 2 * self.batch * 4 + 2 * 2 * self.batch * 16
        self.d2h = (n - 1) * 2 * self.batch * (L * 4 + 16)
        how = f"cb200_fuchs_batch_host: {'one call' if tiles == 1 else str(tiles) + ' calls (instance tiles of ' + str(tb) + ', host arrays reused)'} per step, " \
              f"caller-owned full result arrays ({'page-locked' if pm and pt else 'PAGEABLE'}), " \
              "initial rows and operator uploaded, coefficient ranges copied back on a second stream while the next range is computed"
        return how

    def e2e_step(self):
        tb = self.batch // self.e2e_tiles
        for k in range(self.e2e_tiles):
            self.h_res_m[:2] = self.h_init[k][0]
            self.h_res_t[:2] = self.h_init[k][1]
            rc = self.lib.cb200_fuchs_batch_host(L, ORDER, DEGREE, -2, self.n, tb, self.ode_m, self.ode_t, 1,
                                                 self.h_res_m, self.h_res_t.view(np.int32), self.dev.index)
            assert rc == 0, f"cb200_fuchs_batch_host -> {rc}"

    def e2e_check(self):
        from cascade_b200.format import FLAG_NAN
        t = self.h_res_t
        assert not (t[:, :, 1] & FLAG_NAN).any() and (t[4:, :, 2] != 0).all(), "e2e result rows were not written"


class FrobeniusSweep(Workload):
    cfg = 2
    dtype = "u32 limbs (4096-bit ball)"
    Lh = 128

    def setup(self):
        from cascade_b200.format import instance_major_to_entry_major, pack_entries
        torch, lib, api, dev, args = self.torch, self.lib, self.api, self.dev, self.args
        Lh = self.Lh
        self.batch, self.n = args.batch or 16384, args.terms or 256
        batch, n = self.batch, self.n
        drv = api.HostBufferDriver(dev.index)
        a, b, c = hypgeom_params(batch, Lh, seed=0x5EED0003 + self.rank)
        ode = api.acb_ode_hypgeom_batch(a, b, c, driver=drv)
        from cascade_b200.format import from_complex
        one_minus_c = api.ew(api.OP_SUB, from_complex([1] * batch, Lh), c, driver=drv)
        self.ode_m, self.ode_t = pack_entries(instance_major_to_entry_major(ode, batch, 9), 9, 2 * batch)
        self.rho0_m = np.zeros((1, Lh, 2 * batch), np.uint32)
        self.rho0_t = np.zeros((1, 2 * batch, 4), np.int32)
        self.rho0_t[:, :, 1] = 2
        self.rho1_m, self.rho1_t = pack_entries(one_minus_c, 1, 2 * batch)
        self.merged = getattr(args, "frob_calls", 1) == 1
        if self.merged:
            # ONE call per step: the (operator, exponent) pairs of both runs side by side -- instances 0 .. batch-1 with
            # rho = 0, batch .. 2 batch-1 with rho = 1 - c (the ABI takes an exponent per instance)
            self.ode_m = np.ascontiguousarray(np.concatenate([self.ode_m, self.ode_m], axis=2))
            self.ode_t = np.ascontiguousarray(np.concatenate([self.ode_t, self.ode_t], axis=1))
            self.rho_m = np.ascontiguousarray(np.concatenate([self.rho0_m, self.rho1_m], axis=2))
            self.rho_t = np.ascontiguousarray(np.concatenate([self.rho0_t, self.rho1_t], axis=1))
            self.d_rho = [self.to_dev(self.rho_m, self.rho_t)]
            self.call_batch = 2 * batch
        else:
            self.d_rho = [self.to_dev(self.rho0_m, self.rho0_t), self.to_dev(self.rho1_m, self.rho1_t)]
            self.call_batch = batch
        self.d_ode = self.to_dev(self.ode_m, self.ode_t)
        self.d_res_m = torch.empty((n + 1, Lh, 2 * self.call_batch), dtype=torch.int32, device=dev)
        self.d_res_t = torch.empty((n + 1, 2 * self.call_batch, 4), dtype=torch.int32, device=dev)
        self.ws_bytes = lib.cb200_frobenius_workspace_bytes(Lh, self.call_batch)
        self.d_ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.coeffs = 2 * batch * n
        self.config = self.describe()

    metric = "ball-coeffs/sec, 4096-bit batched acb_ode_solve_frobenius (hypergeometric sweep)"

    def describe(self):
        batch, n = self.args.batch or 16384, self.args.terms or 256
        calls = getattr(self.args, "frob_calls", 1)
        how = "one step = both runs, submitted as ONE batch of (operator, exponent) pairs" if calls == 1 else "one step = both runs, one call each"
        return {"workload": f"configs[2]: Gauss hypergeometric acb_ode_solve_frobenius (M = 1), 4096 bits, {batch} (a,b,c) "
                            f"triples per GPU, {n} terms, rho = 0 and rho = 1 - c ({how})", "calls_per_step": calls,
                "L_limbs": self.Lh, "order": 2, "degree": 2, "terms": n, "batch_per_gpu": batch, "seed": hex(0x5EED0003),
                "cache": "4.4 GB of results and 0.5 GB of per-instance temporaries per run exceed the 126 MB L2; no flush needed"}

    def prepare(self):
        pass

    def kernels(self):
        p = self.p
        for rm, rt in self.d_rho:
            rc = self.lib.cb200_frobenius1_batch_dev(self.Lh, 2, 2, -1, self.n, self.call_batch, p(self.d_ode[0]), p(self.d_ode[1]), 0,
                                                     p(rm), p(rt), 0, p(self.d_res_m), p(self.d_res_t), p(self.d_ws),
                                                     self.ws_bytes, self.sp)
            assert rc == 0, f"cb200_frobenius1_batch_dev -> {rc}"

    def finish(self):
        pass

    def roofline(self, kernel_ms, peaks, imad_peak):
        Lh = self.Lh
        macs = 36 * Lh * Lh * self.coeffs  # 7 complex products + 1 complex division
        return imad_roofline(macs, kernel_ms, imad_peak, self.coeffs * 2 * (4 * Lh + 16), peaks,
                             traffic_for(2, {"batch": self.batch, "terms": self.n, "calls": 1 if self.merged else 2}))

    def e2e_setup(self):
        Lh, n, S = self.Lh, self.n, 2 * self.call_batch
        self.h_res_m, pm = pinned_array(self.lib, (n + 1, Lh, S), np.uint32)
        self.h_res_t, _ = pinned_array(self.lib, (n + 1, S, 4), np.int32)
        del self.d_res_m, self.d_res_t, self.d_ws
        self.torch.cuda.empty_cache()
        calls = 1 if self.merged else 2
        if self.merged:
            self.h2d = self.ode_m.nbytes + self.ode_t.nbytes + self.rho_m.nbytes + self.rho_t.nbytes
        else:
            self.h2d = 2 * (self.ode_m.nbytes + self.ode_t.nbytes) + self.rho0_m.nbytes + self.rho0_t.nbytes + self.rho1_m.nbytes + self.rho1_t.nbytes
        self.d2h = calls * (n + 1) * S * (Lh * 4 + 16)
        return f"cb200_frobenius1_batch_host, {calls} call(s) per step (rho = 0 and rho = 1 - c): operators and exponents uploaded, " \
               f"the full series copied back into caller-owned {'page-locked' if pm else 'pageable'} arrays"

    def e2e_step(self):
        runs = ((self.rho_m, self.rho_t),) if self.merged else ((self.rho0_m, self.rho0_t), (self.rho1_m, self.rho1_t))
        for rm, rt in runs:
            rc = self.lib.cb200_frobenius1_batch_host(self.Lh, 2, 2, -1, self.n, self.call_batch, self.ode_m, self.ode_t, 0, rm, rt, 0,
                                                      self.h_res_m, self.h_res_t, self.dev.index)
            assert rc == 0, f"cb200_frobenius1_batch_host -> {rc}"

    def e2e_check(self):
        from cascade_b200.format import FLAG_NAN
        assert not (self.h_res_t[:, :, 1] & FLAG_NAN).any()


class LegendreSweep(Workload):
    cfg = 3
    dtype = "u32 limbs (1024-bit ball), int64 operator coefficients"
    scaling = "strong"
    TILE = 1 << 17

    def setup(self):
        torch, lib, dev, args = self.torch, self.lib, self.dev, self.args
        self.total = args.batch or (1 << 20)
        self.n = args.terms or 1000
        n = self.n
        from cascade_b200.sharding import partition
        self.lo, self.hi = partition(self.total, self.world, self.rank)
        self.mine = self.hi - self.lo
        self.tile = min(self.TILE, max(self.mine, 1))
        S = 2 * self.tile
        self.d_res_m = torch.zeros((n + 1, L, S), dtype=torch.int32, device=dev)
        self.d_res_t = torch.zeros((n + 1, S, 4), dtype=torch.int32, device=dev)
        self.d_res_t[:, :, 1] = 2
        # c0 = c1 = 1 (both parities of the series), exact
        self.d_res_m[0:2, L - 1, 0::2] = -(2 ** 31)
        self.d_res_t[0:2, 0::2, 0] = 1
        self.d_res_t[0:2, 0::2, 1] = 0
        self.tiles = []
        for t0 in range(self.lo, self.hi, self.tile):
            t1 = min(t0 + self.tile, self.hi)
            ns = torch.arange(t0, t1, dtype=torch.int64, device=dev)
            oi = torch.zeros((9, t1 - t0), dtype=torch.int64, device=dev)
            oi[0] = ns * (ns + 1)      # acb_ode_legendre (reference src/examples.c:3-11): n(n+1), -2z, 1 - z^2
            oi[4] = -2
            oi[6] = 1
            oi[8] = -1
            self.tiles.append((t1 - t0, oi))
        # the reduction that is gathered: the last coefficient of every instance of this rank's shard
        self.nmax = max(partition(self.total, self.world, r)[1] - partition(self.total, self.world, r)[0] for r in range(self.world))
        self.d_last_m = torch.zeros((L, 2 * self.nmax), dtype=torch.int32, device=dev)
        self.d_last_t = torch.zeros((2 * self.nmax, 4), dtype=torch.int32, device=dev)
        if self.world > 1:
            self.g_m = torch.empty((self.world, L, 2 * self.nmax), dtype=torch.int32, device=dev)
            self.g_t = torch.empty((self.world, 2 * self.nmax, 4), dtype=torch.int32, device=dev)
        self.coeffs = self.mine * (n - 1)
        self.config = self.describe()

    metric = "ball-coeffs/sec, 1024-bit Legendre sweep (integer-operator acb_ode_solve_fuchs)"

    def describe(self):
        from cascade_b200.sharding import partition
        total, n = self.args.batch or (1 << 20), self.args.terms or 1000
        lo, hi = partition(total, self.world, 0)
        tile = min(self.TILE, max(hi - lo, 1))
        return {"workload": f"configs[3]: Legendre sweep n = 0 .. {total - 1}, {n} terms, 1024 bits, the whole sweep per step "
                            f"in tiles of {tile} instances per launch, sharded over the ranks (strong scaling), NCCL "
                            "all-gather of the last coefficient of every instance inside the step",
                "L_limbs": L, "order": 2, "degree": 2, "terms": n, "instances_total": total, "tile": tile,
                "cache": f"every tile writes {tile * (n - 1) * 288 / 1e9:.1f} GB of series: far beyond the 126 MB L2; no flush needed",
                "gather": "full series (302 GB at the BASELINE size) stay on the device that computed them; the gathered "
                          "reduction is coefficient n of every instance (302 MB at the BASELINE size)"}

    def prepare(self):
        pass

    def kernels(self):
        p, n = self.p, self.n
        off = 0
        for cnt, oi in self.tiles:
            S = 2 * cnt
            assert cnt == self.tile, "instances per rank must be a multiple of the tile (the result arrays are laid out per tile)"
            rc = self.lib.cb200_fuchs_intop_batch_dev(L, 2, 2, -2, n, cnt, p(oi), 0, p(self.d_res_m), p(self.d_res_t), self.sp)
            assert rc == 0, f"cb200_fuchs_intop_batch_dev -> {rc}"
            self.d_last_m[:, off:off + S].copy_(self.d_res_m[n])
            self.d_last_t[off:off + S].copy_(self.d_res_t[n])
            off += S

    def finish(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.g_m.view(-1), self.d_last_m.view(-1))
            dist.all_gather_into_tensor(self.g_t.view(-1), self.d_last_t.view(-1))

    def roofline(self, kernel_ms, peaks, imad_peak):
        bytes_out = self.coeffs * 2 * (4 * L + 16)  # complex balls as stored (SURVEY counts the real half only: 144 B)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        ach = bytes_out / (kernel_ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": traffic_for(3, {"terms": self.n, "tile": self.tile}), "algorithmic_bytes": bytes_out,
                "algorithmic_bytes_per_coeff": 288, "kernel_ms": kernel_ms,
                "peak_source": "MEASURED_PEAKS.json (copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650",
                "note": "288 B per coefficient written (re + im slots; the imaginary halves of this real sweep are exact zeros "
                        "and are stored like every other ball); on SURVEY 8(d)'s 144 B real-ball convention the fraction halves"}

    def e2e_setup(self):
        n, cnt = self.n, self.tile
        S = 2 * cnt
        self.h_res_m, pm = pinned_array(self.lib, (n + 1, L, S), np.uint32)
        self.h_res_t, _ = pinned_array(self.lib, (n + 1, S, 4), np.int32)
        self.h_res_t[:, :, 1] = 2
        self.h_res_m[0:2, L - 1, 0::2] = 0x80000000
        self.h_res_t[0:2, 0::2, 0] = 1
        self.h_res_t[0:2, 0::2, 1] = 0
        self.h_oi = self.tiles[0][1].cpu().numpy()
        del self.d_res_m, self.d_res_t
        self.torch.cuda.empty_cache()
        self.e2e_coeffs = cnt * (n - 1)
        self.h2d = self.h_oi.nbytes + 2 * S * (4 * L + 16)
        self.d2h = (n + 1) * S * (4 * L + 16)
        return f"cb200_fuchs_intop_batch_host on ONE tile of {cnt} instances per step (the sweep is {len(self.tiles)} such tiles per " \
               f"rank; a host cannot hold the 302 GB of a whole sweep): the tile's full series copied back into caller-owned " \
               f"{'page-locked' if pm else 'pageable'} arrays"

    def e2e_step(self):
        rc = self.lib.cb200_fuchs_intop_batch_host(L, 2, 2, -2, self.n, self.tile, self.h_oi, 0, self.h_res_m, self.h_res_t,
                                                   self.dev.index)
        assert rc == 0, f"cb200_fuchs_intop_batch_host -> {rc}"

    def e2e_check(self):
        from cascade_b200.format import FLAG_NAN
        assert not (self.h_res_t[:, :, 1] & FLAG_NAN).any()


class HornerPoints(Workload):
    cfg = 4
    dtype = "u32 limbs (2048-bit ball)"
    Lh = 64

    def setup(self):
        from cascade_b200.format import pack_entries, random_balls
        torch, lib, dev, args = self.torch, self.lib, self.dev, self.args
        Lh = self.Lh
        self.npts, self.length = args.batch or 10 ** 6, args.terms or 3504
        rng = np.random.default_rng(0x5EED0005)
        f = random_balls(rng, 2 * self.length, Lh, -1, 1, "zero")       # a dense series, as a continued solution is
        pts = random_balls(np.random.default_rng(0x5EED0005 + 1 + self.rank), 2 * self.npts, Lh, -3, -2, "zero")  # |x| <= 1/4
        self.f_m, self.f_t = pack_entries(f, self.length, 2)
        self.p_m, self.p_t = pack_entries(pts, 1, 2 * self.npts)
        self.d_f = self.to_dev(self.f_m, self.f_t)
        self.d_p = self.to_dev(self.p_m, self.p_t)
        self.d_o_m = torch.empty((1, Lh, 2 * self.npts), dtype=torch.int32, device=dev)
        self.d_o_t = torch.empty((1, 2 * self.npts, 4), dtype=torch.int32, device=dev)
        self.ws_bytes = lib.cb200_horner_workspace_bytes(Lh, self.npts)
        self.d_ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.coeffs = self.npts * (self.length - 1)
        self.config = self.describe()

    metric = "ball-coeffs/sec, 2048-bit Horner evaluation (series coefficients consumed x points)"

    def describe(self):
        npts, length = self.args.batch or 10 ** 6, self.args.terms or 3504
        return {"workload": f"configs[4]: Horner evaluation of one {length}-term series (truncation_order of the 64-step "
                            f"continuation path at 2048 bits) at {npts} points per GPU, 2048 bits",
                "L_limbs": self.Lh, "terms": length, "points_per_gpu": npts, "seed": hex(0x5EED0005),
                "cache": "the series (1.9 MB) is meant to stay in L2; the points and accumulators (0.5 GB + temporaries) "
                         "exceed it; no flush needed",
                "continuation_leg": "timed by tools/bench_configs.py --continuation (sequential in the path: not a bench step)"}

    def prepare(self):
        pass

    def kernels(self):
        p = self.p
        rc = self.lib.cb200_horner_batch_dev(self.Lh, self.length, self.npts, 1, p(self.d_f[0]), p(self.d_f[1]), 1, p(self.d_p[0]),
                                             p(self.d_p[1]), p(self.d_o_m), p(self.d_o_t), p(self.d_ws), self.ws_bytes, self.sp)
        assert rc == 0, f"cb200_horner_batch_dev -> {rc}"

    def finish(self):
        pass

    def roofline(self, kernel_ms, peaks, imad_peak):
        Lh = self.Lh
        macs = 4 * Lh * Lh * self.coeffs  # one complex multiplication per series term and point
        return imad_roofline(macs, kernel_ms, imad_peak, self.npts * 2 * (4 * Lh + 16), peaks,
                             traffic_for(4, {"terms": self.length, "points": self.npts}))

    def e2e_setup(self):
        Lh = self.Lh
        self.h_o_m, pm = pinned_array(self.lib, (1, Lh, 2 * self.npts), np.uint32)
        self.h_o_t, _ = pinned_array(self.lib, (1, 2 * self.npts, 4), np.int32)
        self.h2d = self.f_m.nbytes + self.f_t.nbytes + self.p_m.nbytes + self.p_t.nbytes
        self.d2h = self.h_o_m.nbytes + self.h_o_t.nbytes
        return "cb200_horner_batch_host: series and points uploaded, values copied back"

    def e2e_step(self):
        rc = self.lib.cb200_horner_batch_host(self.Lh, self.length, self.npts, 1, self.f_m, self.f_t, 1, self.p_m, self.p_t,
                                              self.h_o_m, self.h_o_t, self.dev.index)
        assert rc == 0, f"cb200_horner_batch_host -> {rc}"

    def e2e_check(self):
        from cascade_b200.format import FLAG_NAN
        assert not (self.h_o_t[:, :, 1] & FLAG_NAN).any()


WORKLOADS = {1: FuchsShared, 2: FrobeniusSweep, 3: LegendreSweep, 4: HornerPoints}


def describe(cfg, args, world=1):
    """The `config` object of a workload without touching a GPU (the reference arm prints the same one)."""
    class _NoGpu:
        pass
    w = WORKLOADS[cfg].__new__(WORKLOADS[cfg])
    w.args, w.rank, w.world = args, 0, world
    return w.describe()


def traffic_for(cfg, key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of this exact config
    (profiles/traffic_r02.json, written by tools/ncu_summary.py from the raw ncu export); None when the sizes differ."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))[str(cfg)]
    except Exception:
        return None
    if all(rec.get("config", {}).get(k) == v for k, v in key.items()):
        return rec.get("dram_bytes_per_step")
    return None


def imad_roofline(macs, kernel_ms, imad_peak, bytes_out, peaks, traffic):
    achieved = macs / (kernel_ms * 1e-3) / 1e12
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    return {"bound": "imad", "achieved": achieved, "peak": imad_peak, "unit": "T limb-MAC/s", "frac": achieved / imad_peak,
            "traffic": traffic, "algorithmic_bytes": bytes_out,
            "peak_source": "measured plain IMAD.WIDE.U32 rate, profiles/imad_peak_r01.json (carry-chained IMAD.WIDE.X peaks at 8.55)",
            "kernel_ms": kernel_ms,
            "hbm_view": {"achieved": bytes_out / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": bytes_out / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                         "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4], help="index into BASELINE.json configs (default 1: the headline)")
    ap.add_argument("--batch", type=int, default=0, help="instances (points) per GPU -- config 3: instances of the whole sweep (default: the BASELINE size)")
    ap.add_argument("--terms", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-tiles", type=int, default=0, help="config 1: force the number of instance tiles of the end-to-end leg (default: 1 if the host memory allows)")
    ap.add_argument("--frob-calls", type=int, default=1, choices=[1, 2], help="config 2: both exponents as one batch (1) or one call per exponent (2)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    # exactly ONE line on stdout: everything libraries print (NCCL's version banner, warnings) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    cfg = args.config

    if args.impl == "reference":
        if rank != 0:
            return
        W = WORKLOADS[cfg]
        vals, sample, dts = [], "", []
        for _ in range(max(1, args.steps)):
            t0 = time.perf_counter()
            v, sample = cpu_port(cfg, ncores, 8.0, args)
            dts.append(time.perf_counter() - t0)
            vals.append(v)
        val = sum(vals) / len(vals)
        line = {"impl": "reference", "metric": W.metric, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(dts) / len(dts) * 1e3,
                "higher_is_better": True, "scaling": W.scaling, "vs_baseline": None, "dtype": W.dtype,
                "data": "synthetic", "config": describe(cfg, args, args.gpus),
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample + " per step"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from cascade_b200 import _lib, api

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    w = WORKLOADS[cfg](lib, api, torch, dev, rank, world, args)
    w.setup()
    stream = w.stream

    def step_device():
        w.prepare()
        w.kernels()
        w.finish()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local_rank), daemon=True)
    for i in range(args.warmup):
        if i == args.warmup - 1:
            th.start()  # sampler start-up (nvidia-smi spawn) happens during the last warm-up step
        step_device()
    if not th.is_alive():
        th.start()
    barrier()
    samples.clear()
    launches0 = lib.cb200_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0.record(stream)
    for k in range(args.steps):
        w.prepare()
        kev[k][0].record(stream)
        w.kernels()
        kev[k][1].record(stream)
        w.finish()
    e1.record(stream)
    barrier()
    stop.set()
    th.join(timeout=2)
    launches = lib.cb200_kernel_launches() - launches0
    ms_total = e0.elapsed_time(e1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    tc = torch.tensor([float(w.coeffs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.SUM)
    value = float(tc.item()) / (ms_step * 1e-3)

    # configs[1]: the only collective of the path is a gather of results; the full series (37.7 GB per GPU) stay
    # sharded, the last coefficient of every instance is gathered (outside the timed region: not part of the solve)
    if world > 1 and cfg == 1:
        last = w.d_res_m[w.n].contiguous()
        gathered = [torch.empty_like(last) for _ in range(world)]
        dist.all_gather(gathered, last)
        torch.cuda.synchronize(dev)

    # ---- end to end: through the host-buffer C-ABI call, host arrays in and out, copies inside the timing
    e2e = None
    e2e_note = None
    if not args.no_e2e:
        # the caller-owned result arrays of every rank of this box must fit the host's memory (configs[1]: 37.7 GB per
        # rank); if they do not, the end-to-end leg is skipped rather than measured on something smaller
        need = getattr(w, "e2e_host_bytes", lambda: 0)() * world
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = None
        tiles = max(1, args.e2e_tiles)
        while avail is not None and need / tiles * 1.25 > avail and tiles < 16 and cfg == 1:
            tiles *= 2
        short = avail is not None and need / tiles * 1.25 > avail
        flag = torch.tensor([tiles, 1 if short else 0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        tiles = int(flag[0].item())
        if int(flag[1].item()):
            e2e_note = f"skipped: {need / 1e9:.0f} GB of host result arrays for {world} rank(s) against {0 if avail is None else avail / 1e9:.0f} GB available"
    if not args.no_e2e and e2e_note is None:
        how = w.e2e_setup(tiles) if cfg == 1 else w.e2e_setup()
        w.e2e_step()
        barrier()
        e2e_steps = max(1, min(args.steps, 2))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            w.e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        w.e2e_check()
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ecoeffs = torch.tensor([float(getattr(w, "e2e_coeffs", w.coeffs))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ecoeffs, op=dist.ReduceOp.SUM)
        e2e = {"value": float(ecoeffs.item()) / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": int(w.h2d) * world,
               "d2h_bytes_per_step": int(w.d2h) * world, "ms_per_step": float(tt.item()) * 1e3, "how": how}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    imad_peak = 18.44
    try:
        imad_peak = float(json.loads(open(os.path.join(ROOT, "profiles", "imad_peak_r01.json")).read())["imad_wide_plain_Tops"])
    except Exception:
        pass
    roofline = w.roofline(kernel_ms, peaks, imad_peak)
    cpu = None
    if not args.no_cpu:
        v, sample = cpu_port(cfg, ncores, 10.0, args)
        v1, sample1 = cpu_port(cfg, 1, 3.0, args)  # the reference is single-threaded as shipped
        cpu = {"value": v, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample,
               "single_core": {"value": v1, "sample": sample1}}
    line = {"metric": w.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": w.scaling,
            "vs_baseline": None, "dtype": w.dtype, "data": "synthetic", "config": w.config,
            "clocks": summarise_clocks(samples), "e2e": e2e if e2e is not None else ({"skipped": e2e_note} if e2e_note else None),
            "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
