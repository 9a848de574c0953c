#!/usr/bin/env python
"""bench.py -- throughput of the batched acb_ode_solve_fuchs hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    (N > 1: launched by torchrun, one rank per GPU)

Workload = BASELINE.json configs[1]: Bessel ODE (complex nu) shifted to a complex ordinary point,
2000-term Fuchs series at 1024 bits, 65536 initial-condition pairs per GPU, one operator shared
by the batch.  A "step" is one full solve of the batch.  Metric: ball-coefficients per second
(complex series coefficients written / time).

  value : device-resident (inputs already in HBM, CUDA events around the kernels)
  e2e   : through the host-buffer path -- pinned host inputs copied H2D and the full result
          series copied D2H inside the timed region, one coefficient range after the other while
          the next range is being computed
  roofline : integer pipe.  achieved = 24 L^2 limb multiply-adds per coefficient
          x coefficients / kernel time; peak = measured plain IMAD.WIDE rate
          (profiles/imad_peak_r01.json).  The HBM view (bytes written / time vs the measured copy
          bandwidth) is reported next to it.
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

L = 32            # 1024 bits
ORDER = DEGREE = 2
N_TERMS = 2000    # coefficients 0..2000; 1999 new ones per instance
BATCH = 65536     # per GPU
SEED = 0x5EED0002
METRIC = "ball-coeffs/sec, 1024-bit batched acb_ode_solve_fuchs"
UNIT = "ball-coeffs/s"


def make_inputs(batch: int, n: int, drv_for_setup=None):
    """Bessel(nu) shifted to a, and `batch` random initial-value pairs (host Balls)."""
    from cascade_b200 import api
    from cascade_b200.format import random_balls
    rng = np.random.default_rng(SEED)
    nu = random_balls(rng, 2, L, 0, 0, "zero")
    a = random_balls(rng, 2, L, 0, 1, "zero")
    ode0, _ = api.acb_ode_bessel_batch(nu, driver=drv_for_setup)
    ode = api.acb_ode_shift_batch(ode0, a, ORDER, DEGREE, driver=drv_for_setup)
    init = random_balls(rng, batch * 2 * 2, L, -1, 0, "zero")
    return ode, init


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


def cpu_port(ncores: int, target_seconds: float = 10.0):
    """Time the oracle port on the host: same operator, n = N_TERMS, a bounded number of instances."""
    from oracle import cbo
    from cascade_b200.format import random_balls
    rng = np.random.default_rng(SEED)
    nu = random_balls(rng, 2, L, 0, 0, "zero")
    a = random_balls(rng, 2, L, 0, 1, "zero")
    ode0 = cbo.example("bessel", L, 0, nu)
    ode = cbo.ode_shift(ORDER, DEGREE, ode0, a)
    # calibrate on one instance per core, then size the sample
    def run(batch):
        init = random_balls(rng, batch * 4, L, -1, 0, "zero")
        from oracle.cbo import _series_with_init, lib
        res = _series_with_init(init, batch, N_TERMS, 2)
        t0 = time.perf_counter()
        rc = lib().cbo_fuchs_batch(L, ORDER, DEGREE, N_TERMS, batch, 1, ode.mant, ode.meta, res.mant, res.meta, ncores)
        dt = time.perf_counter() - t0
        assert rc == 0
        return dt
    dt1 = run(ncores)
    per_inst = dt1 / 1.0  # seconds per "one instance on every core"
    rounds = max(1, min(64, int(target_seconds / max(per_inst, 1e-3))))
    batch = ncores * rounds
    dt = run(batch)
    coeffs = batch * (N_TERMS - 1)
    return coeffs / dt, batch, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=BATCH, help="instances per GPU (default: the BASELINE config)")
    ap.add_argument("--terms", type=int, default=N_TERMS)
    ap.add_argument("--no-e2e", action="store_true")
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
    config = {"workload": "configs[1]: Bessel ODE (complex nu) shifted to an ordinary point, "
                          f"{args.terms}-term Fuchs series, 1024 bits, {args.batch} initial conditions per GPU, shared operator",
              "L_limbs": L, "order": ORDER, "degree": DEGREE, "terms": args.terms, "batch_per_gpu": args.batch,
              "seed": hex(SEED),
              "cache": "outputs (tens of GB per step) and the coefficient window far exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        v, sample_batch, dt = cpu_port(ncores, target_seconds=8.0)
        per = []
        for _ in range(max(0, args.steps - 1)):
            per.append(cpu_port(ncores, target_seconds=8.0)[0])
        vals = [v] + per
        val = sum(vals) / len(vals)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (1024-bit ball)",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": "port",
                                 "sample": f"{sample_batch} instances x {args.terms - 1} coefficients per step, "
                                           "oracle port (Cascade loops on GMP 6.3 mpn); Cascade+Arb itself is not buildable here"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from cascade_b200 import _lib, api
    from cascade_b200.format import instance_major_to_entry_major, pack_entries

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    drv = api.HostBufferDriver(local_rank)
    batch, n = args.batch, args.terms
    S = 2 * batch

    # ---- inputs (host), then resident device arrays
    ode, init = make_inputs(batch, n, drv)
    # different initial values on every rank (independent instances; weak scaling)
    if rank:
        from cascade_b200.format import random_balls
        init = random_balls(np.random.default_rng(SEED + rank), batch * 4, L, -1, 0, "zero")
    ode_m, ode_t = api._ode_to_device(ode, ORDER, DEGREE, batch, True)
    init_m, init_t = pack_entries(instance_major_to_entry_major(init, batch, 2), 2, S)
    d_ode_m = torch.from_numpy(ode_m.view(np.int32)).to(dev)
    d_ode_t = torch.from_numpy(ode_t).to(dev)
    d_init_m = torch.from_numpy(init_m.view(np.int32)).to(dev)
    d_init_t = torch.from_numpy(init_t).to(dev)
    d_res_m = torch.empty((n + 1, L, S), dtype=torch.int32, device=dev)
    d_res_t = torch.empty((n + 1, S, 4), dtype=torch.int32, device=dev)
    ws_bytes = lib.cb200_fuchs_workspace_bytes(L, 2, 2, -2, n, batch, 1)
    d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    stream = torch.cuda.current_stream(dev)
    sp = C.c_void_p(stream.cuda_stream)

    def step_device():
        d_res_m[:2].copy_(d_init_m)
        d_res_t[:2].copy_(d_init_t)
        rc = lib.cb200_fuchs_batch_dev(L, ORDER, DEGREE, -2, n, batch, p(d_ode_m), p(d_ode_t), 1, p(d_res_m),
                                       p(d_res_t), p(d_ws), ws_bytes, sp)
        assert rc == 0, f"cb200_fuchs_batch_dev -> {rc}"

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
        d_res_m[:2].copy_(d_init_m)
        d_res_t[:2].copy_(d_init_t)
        kev[k][0].record(stream)
        rc = lib.cb200_fuchs_batch_dev(L, ORDER, DEGREE, -2, n, batch, p(d_ode_m), p(d_ode_t), 1, p(d_res_m),
                                       p(d_res_t), p(d_ws), ws_bytes, sp)
        kev[k][1].record(stream)
        assert rc == 0
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
    coeffs_per_gpu = batch * (n - 1)
    value = world * coeffs_per_gpu / (ms_step * 1e-3)

    # the only collective of the path: gather the last coefficient of every instance (NVLink)
    if world > 1:
        last = d_res_m[n].contiguous()
        gathered = [torch.empty_like(last) for _ in range(world)]
        dist.all_gather(gathered, last)
        torch.cuda.synchronize(dev)

    # ---- end to end: host buffers in, host buffers out, tile by tile, copies inside the timing
    e2e = None
    if not args.no_e2e:
        # The solve runs over the whole batch in coefficient RANGES (cb200_fuchs_batch_range_dev: the recurrence
        # resumes from the rows already in HBM); the finished rows of a range -- contiguous in the
        # [coefficient][limb][slot] layout -- go to pinned host memory on a second stream while the next range is
        # being computed.  Pinned ring of two range buffers per rank, kept within a quarter of the free host memory
        # across the ranks of the box (a real caller would own a full-size result buffer).
        row_bytes = S * (L * 4 + 16)
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 64 << 30
        rows = 250
        while rows > 10 and 2 * rows * row_bytes * world > avail // 4:
            rows //= 2
        bounds = list(range(0, n + 1, rows)) + [n + 1]
        nchunks = len(bounds) - 1
        h_init_m = torch.from_numpy(init_m.view(np.int32)).pin_memory()
        h_init_t = torch.from_numpy(init_t).pin_memory()
        h_res_m = torch.empty((2, rows, L, S), dtype=torch.int32).pin_memory()
        h_res_t = torch.empty((2, rows, S, 4), dtype=torch.int32).pin_memory()
        s_comp, s_copy = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        ev = [torch.cuda.Event() for _ in range(nchunks)]
        h2d = 2 * L * S * 4 + 2 * S * 16
        d2h = (n + 1) * row_bytes

        def e2e_step():
            with torch.cuda.stream(s_comp):
                d_res_m[:2].copy_(h_init_m, non_blocking=True)
                d_res_t[:2].copy_(h_init_t, non_blocking=True)
            for c in range(nchunks):
                lo, hi = bounds[c], bounds[c + 1]
                with torch.cuda.stream(s_comp):
                    rc = lib.cb200_fuchs_batch_range_dev(L, ORDER, DEGREE, -2, lo, hi - 1, batch, p(d_ode_m), p(d_ode_t), 1,
                                                         p(d_res_m), p(d_res_t), p(d_ws), ws_bytes,
                                                         C.c_void_p(s_comp.cuda_stream))
                    assert rc == 0
                    ev[c].record(s_comp)
                with torch.cuda.stream(s_copy):
                    s_copy.wait_event(ev[c])
                    h_res_m[c % 2, : hi - lo].copy_(d_res_m[lo:hi], non_blocking=True)
                    h_res_t[c % 2, : hi - lo].copy_(d_res_t[lo:hi], non_blocking=True)
            s_comp.synchronize()
            s_copy.synchronize()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 2))
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * coeffs_per_gpu / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
               "how": f"{nchunks} coefficient ranges of {rows} rows over the whole batch, compute and copy streams, "
                      "pinned host buffers, the full result series copied back"}

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
    macs = 24 * L * L * coeffs_per_gpu  # (4W+8) L^2 per coefficient, W = 4
    achieved = macs / (kernel_ms * 1e-3) / 1e12
    bytes_out = coeffs_per_gpu * 2 * (4 * L + 16)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # DRAM bytes per launch of the recurrence kernel, ncu on this exact config (profiles/ncu_fuchs_shared_bench_r01.txt:
    # persistent kernel, 10.2 GB read + 41.2 GB written: work units migrate between SMs, so part of the 4-coefficient
    # window comes back from DRAM instead of L2); null for other sizes
    traffic = 51.47e9 if (batch == BATCH and n == N_TERMS) else None
    roofline = {"bound": "imad", "achieved": achieved, "peak": imad_peak, "unit": "T limb-MAC/s",
                "frac": achieved / imad_peak, "traffic": traffic, "algorithmic_bytes": bytes_out,
                "peak_source": "measured plain IMAD.WIDE.U32 rate, profiles/imad_peak_r01.json "
                               "(carry-chained IMAD.WIDE.X peaks at 8.55)",
                "kernel_ms": kernel_ms,
                "hbm_view": {"achieved": bytes_out / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": bytes_out / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}}
    cpu = None
    if not args.no_cpu:
        v, sb, dt = cpu_port(ncores, target_seconds=10.0)
        v1, sb1, dt1 = cpu_port(1, target_seconds=3.0)  # the reference is single-threaded as shipped
        cpu = {"value": v, "unit": UNIT, "cores": ncores, "kind": "port",
               "sample": f"{sb} instances x {n - 1} coefficients, oracle port (Cascade loops on GMP 6.3 mpn), {dt:.1f} s",
               "single_core": {"value": v1, "sample": f"{sb1} instances, {dt1:.1f} s"}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 limbs (1024-bit ball)", "data": "synthetic", "config": config,
            "clocks": summarise_clocks(samples), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
