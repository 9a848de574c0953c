"""Quick device-side timing of the elementwise ball operations (development aid): cost of acb_mul / acb_div / acb_inv."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from cascade_b200 import _lib

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 224 * 8
lib = _lib.lib()
lib.cb200_ew_workspace_bytes.restype = C.c_size_t
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(1)
S = 2 * n


def rand_arr():
    m = torch.randint(0, 2**31 - 1, (1, L, S), device=dev, dtype=torch.int32, generator=g) * 2 + 1
    m[:, L - 1, :] |= -(2**31)
    t = torch.zeros((1, S, 4), device=dev, dtype=torch.int32)
    return m, t


p = lambda t: C.c_void_p(t.data_ptr())
x_m, x_t = rand_arr()
y_m, y_t = rand_arr()
z_m, z_t = rand_arr()
wsb = lib.cb200_ew_workspace_bytes(L, n)
ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
k = torch.full((n,), 12345, device=dev, dtype=torch.int64)
for name, op in (("add", 0), ("mul", 2), ("div", 3), ("inv", 4), ("mul_si", 16)):
    for it in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.cb200_ew_dev(op, L, n, p(z_m), p(z_t), p(x_m), p(x_t), p(y_m), p(y_t), p(k), p(ws), wsb, None)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0, rc
    ms = e0.elapsed_time(e1)
    print(f"L={L} n={n} {name}: {ms:.3f} ms, {n / ms * 1e3:.3e} complex ops/s")
