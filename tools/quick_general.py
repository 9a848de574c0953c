"""Quick timing of the general (per-instance operator) Fuchs kernel and the Frobenius kernel."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from cascade_b200 import _lib

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 33152
n = int(sys.argv[3]) if len(sys.argv) > 3 else 60
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


p = lambda t: C.c_void_p(t.data_ptr())
ode_m, ode_t = rand_arr(9, S)
res_m, res_t = rand_arr(n + 1, S)
wsb = lib.cb200_fuchs_workspace_bytes(L, 2, 2, -2, n, batch, 0)
ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
for it in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.cb200_fuchs_batch_dev(L, 2, 2, -2, n, batch, p(ode_m), p(ode_t), 0, p(res_m), p(res_t), p(ws), wsb, None)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0
    ms = e0.elapsed_time(e1)
    c = batch * (n - 1)
    print(f"fuchs general L={L} batch={batch} n={n}: {ms:.2f} ms, {c / ms * 1e3:.3e} coeff/s, {c * 24 * L * L / ms * 1e3 / 1e12:.3f} T (24L^2)")
rho_m, rho_t = rand_arr(1, S)
wsb = lib.cb200_frobenius_workspace_bytes(L, batch)
ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
for it in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.cb200_frobenius1_batch_dev(L, 2, 2, -1, n, batch, p(ode_m), p(ode_t), 0, p(rho_m), p(rho_t), 0, p(res_m), p(res_t), p(ws), wsb, None)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0
    ms = e0.elapsed_time(e1)
    c = batch * n
    print(f"frobenius L={L} batch={batch} n={n}: {ms:.2f} ms, {c / ms * 1e3:.3e} coeff/s, {c * 36 * L * L / ms * 1e3 / 1e12:.3f} T (36L^2)")
