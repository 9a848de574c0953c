"""Quick device-side timing of the Horner kernel (development aid)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from cascade_b200 import _lib

L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 28416
length = int(sys.argv[3]) if len(sys.argv) > 3 else 200
lib = _lib.lib()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(1)


def rand_arr(entries, S_):
    m = torch.randint(0, 2**31 - 1, (entries, L, S_), device=dev, dtype=torch.int32, generator=g) * 2 + 1
    m[:, L - 1, :] |= -(2**31)
    t = torch.zeros((entries, S_, 4), device=dev, dtype=torch.int32)
    return m, t


f_m, f_t = rand_arr(length, 2)
p_m, p_t = rand_arr(1, 2 * npts)
p_t[:, :, 0] = -2
o_m = torch.empty((1, L, 2 * npts), dtype=torch.int32, device=dev)
o_t = torch.empty((1, 2 * npts, 4), dtype=torch.int32, device=dev)
wsb = lib.cb200_horner_workspace_bytes(L, npts)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
p = lambda t: C.c_void_p(t.data_ptr())
for it in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.cb200_horner_batch_dev(L, length, npts, 1, p(f_m), p(f_t), 1, p(p_m), p(p_t), p(o_m), p(o_t), p(ws), wsb, None)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0
    ms = e0.elapsed_time(e1)
    steps = npts * (length - 1)
    print(f"L={L} npts={npts} len={length}: {ms:.2f} ms, {steps / ms * 1e3:.3e} steps/s, {steps * 4 * L * L / ms * 1e3 / 1e12:.3f} T limb-MAC/s (4L^2)")
