"""Development aid: cooperative (four lanes per instance) vs one-thread kernels at small batches, 2048 / 4096 bits."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from cascade_b200 import _lib

lib = _lib.lib()
dev = torch.device("cuda:0")
p = lambda t: C.c_void_p(t.data_ptr())


def rand_dev(entries, L, S, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    m = torch.randint(0, 2**31 - 1, (entries, L, S), device=dev, dtype=torch.int32, generator=g) * 2 + 1
    m[:, L - 1, :] |= -(2**31)
    t = torch.zeros((entries, S, 4), device=dev, dtype=torch.int32)
    return m, t


def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc = fn(); e1.record(); torch.cuda.synchronize()
    assert rc == 0, rc
    return e0.elapsed_time(e1)


for L in (64, 128):
    length = 200 if L == 64 else 100
    for npts in (2, 64, 1024, 4096, 16384, 65536, 262144):
        f_m, f_t = rand_dev(length, L, 2, 5)
        pts_m, pts_t = rand_dev(1, L, 2 * npts, 6)
        pts_t[:, :, 0] = -2
        o_m = torch.empty((1, L, 2 * npts), dtype=torch.int32, device=dev)
        o_t = torch.empty((1, 2 * npts, 4), dtype=torch.int32, device=dev)
        wsb = lib.cb200_horner_workspace_bytes(L, npts)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        res = []
        for coop in ("0", "1", "2"):
            os.environ["CB200_COOP"] = coop
            res.append(timed(lambda: lib.cb200_horner_batch_dev(L, length, npts, 1, p(f_m), p(f_t), 1, p(pts_m), p(pts_t), p(o_m), p(o_t), p(ws), wsb, None)))
        print(f"horner L={L} len={length} npts={npts}: one-thread {res[0]:.2f} ms, four lanes {res[1]:.2f} ms, two lanes {res[2]:.2f} ms", flush=True)
    n = 24
    for batch in (2, 64, 1024, 4096, 16384):
        om, ot = rand_dev(9, L, 2 * batch, 21)
        rm, rt = rand_dev(1, L, 2 * batch, 22)
        res_m = torch.empty((n + 1, L, 2 * batch), dtype=torch.int32, device=dev)
        res_t = torch.empty((n + 1, 2 * batch, 4), dtype=torch.int32, device=dev)
        wsb = lib.cb200_frobenius_workspace_bytes(L, batch)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        res = []
        for coop in ("0", "4", "2"):
            os.environ["CB200_COOP"] = coop
            res.append(timed(lambda: lib.cb200_frobenius1_batch_dev(L, 2, 2, -2, n, batch, p(om), p(ot), 0, p(rm), p(rt), 0, p(res_m), p(res_t), p(ws), wsb, None)))
        print(f"frobenius L={L} n={n} batch={batch}: one-thread {res[0]:.2f} ms, four lanes {res[1]:.2f} ms, two lanes {res[2]:.2f} ms", flush=True)
