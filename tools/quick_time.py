"""Quick device-side timing of the Fuchs batch kernel (development aid)."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from cascade_b200 import _lib

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lib = _lib.lib()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(1)
S = 2 * batch


def rand_arr(entries, S_):
    m = torch.randint(0, 2**31 - 1, (entries, L, S_), device=dev, dtype=torch.int32, generator=g) * 2 + 1
    m[:, L - 1, :] |= -(2**31)
    t = torch.zeros((entries, S_, 4), device=dev, dtype=torch.int32)
    return m, t


ode_m, ode_t = rand_arr(9, 2)
res_m, res_t = rand_arr(n + 1, S)
ws_bytes = lib.cb200_fuchs_workspace_bytes(L, 2, 2, -2, n, batch, 1)
ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
p = lambda t: C.c_void_p(t.data_ptr())
for it in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.cb200_fuchs_batch_dev(L, 2, 2, -2, n, batch, p(ode_m), p(ode_t), 1, p(res_m), p(res_t), p(ws), ws_bytes, None)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0, rc
    ms = e0.elapsed_time(e1)
    coeffs = batch * (n - 1)
    print(f"L={L} batch={batch} n={n}: {ms:.2f} ms, {coeffs / ms * 1e3:.3e} coeff/s, "
          f"{coeffs * 24 * L * L / ms * 1e3 / 1e12:.3f} T limb-MAC/s (24L^2 convention)")

st = (C.c_uint64 * 2)()
lib.cb200_debug_rr_stats(st)
print("rr accepted / rejected:", st[0], st[1])
