"""Host-side mirror of Cascade's operator interface for the hot path, batched.

The functions keep the reference's names and argument meaning (reference ``src/cascade.h:13,25``,
``src/acb_ode.h:25-41``) with a leading batch dimension; all arithmetic happens in the CUDA
library behind the C ABI of ``include/cascade_b200.h`` -- this module only marshals.

Array convention at THIS level (matches the oracle's, so parity tests compare 1:1): ball arrays
are :class:`cascade_b200.format.Balls` in *instance-major* order, ``[batch][entries]`` complex
balls; the device wants ``[entry][limb][2*batch]`` and the conversion happens here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .format import (FLAG_NAN, FLAG_ZERO, Balls, entry_major_to_instance_major,
                     instance_major_to_entry_major, pack_entries, unpack_entries)

OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_INV = 0, 1, 2, 3, 4
OP_MUL_SI, OP_ADD_SI = 16, 17
UNDEFINED = -0xFFFF  # reference src/acb_ode.c:3


class HostBufferDriver:
    """Calls the ``*_host`` entry points: host buffers in, host buffers out, copies included."""

    def __init__(self, device: int = 0):
        self.device = device

    def fuchs(self, L, order, degree, valuation, n, batch, ode_m, ode_t, shared, res_m, res_t):
        return _lib.lib().cb200_fuchs_batch_host(L, order, degree, valuation, n, batch, ode_m, ode_t,
                                                 int(shared), res_m, res_t, self.device)

    def frobenius1(self, L, order, degree, valuation, n, batch, ode_m, ode_t, shared, rho_m, rho_t,
                   shared_rho, res_m, res_t):
        return _lib.lib().cb200_frobenius1_batch_host(L, order, degree, valuation, n, batch, ode_m, ode_t,
                                                      int(shared), rho_m, rho_t, int(shared_rho), res_m,
                                                      res_t, self.device)

    def horner(self, L, length, npts, nder, f_m, f_t, shared_f, p_m, p_t, o_m, o_t):
        return _lib.lib().cb200_horner_batch_host(L, length, npts, nder, f_m, f_t, int(shared_f), p_m, p_t,
                                                  o_m, o_t, self.device)

    def ew(self, op, L, n, z_m, z_t, x_m, x_t, y_m, y_t, k):
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        return _lib.lib().cb200_ew_host(op, L, n, z_m, z_t, x_m, x_t, vp(y_m), vp(y_t), vp(k), self.device)


_default_driver = None


def default_driver():
    global _default_driver
    if _default_driver is None:
        _default_driver = HostBufferDriver(0)
    return _default_driver


# ---------------------------------------------------------------- operator helpers (host logic)
def acb_ode_valuation(ode: Balls, order: int, degree: int, inst: int = 0) -> int:
    """min_i (ord_z p_i - i): reference ``acb_ode_valuation``, src/acb_ode.c:159-181.  Only exact
    zeros (midpoint 0, radius 0) count, as with ``acb_is_zero``."""
    ne = (order + 1) * (degree + 1)
    meta = ode.meta.reshape(-1, ne, 2, 4)[inst]
    exact_zero = ((meta[..., 1] & FLAG_ZERO) != 0) & (meta[..., 2] == 0)
    zero = exact_zero.all(axis=-1).reshape(order + 1, degree + 1)
    val = degree
    for i in range(order + 1):
        z = 0
        while z <= degree and zero[i, z]:
            z += 1
        val = min(val, z - i)
    return val


def _ode_to_device(ode: Balls, order: int, degree: int, batch: int, shared: bool):
    ne = (order + 1) * (degree + 1)
    nb = 1 if shared else batch
    assert ode.n_slots == nb * ne * 2, "operator array has the wrong size"
    return pack_entries(instance_major_to_entry_major(ode, nb, ne), ne, 2 * nb)


# ---------------------------------------------------------------- solvers
def acb_ode_solve_fuchs_batch(ode: Balls, init: Balls, order: int, degree: int, n: int, batch: int,
                              shared_ode: bool = True, valuation: int | None = None, driver=None) -> Balls:
    """Batched ``acb_ode_solve_fuchs(res, ODE, n, 32*L)`` (reference src/fuchs_solver.c:96-145).

    ``init``: ``[batch][-valuation]`` complex balls (``res[0..-v-1]`` of each call).
    Returns ``[batch][n+1]`` complex balls.  Raises for ``valuation >= 0`` (undefined in the
    reference)."""
    drv = driver or default_driver()
    L = ode.L
    v = acb_ode_valuation(ode, order, degree) if valuation is None else valuation
    if v >= 0:
        _lib.check(_lib.EVALUATION, "acb_ode_solve_fuchs_batch")
    ninit = -v
    assert init.n_slots == batch * ninit * 2, "need -valuation initial values per instance"
    assert n >= ninit - 1
    ode_m, ode_t = _ode_to_device(ode, order, degree, batch, shared_ode)
    res = Balls.zeros(batch * (n + 1) * 2, L)
    res.mant.reshape(batch, n + 1, 2, L)[:, :ninit] = init.mant.reshape(batch, ninit, 2, L)
    res.meta.reshape(batch, n + 1, 2, 4)[:, :ninit] = init.meta.reshape(batch, ninit, 2, 4)
    res_m, res_t = pack_entries(instance_major_to_entry_major(res, batch, n + 1), n + 1, 2 * batch)
    _lib.check(drv.fuchs(L, order, degree, v, n, batch, ode_m, ode_t, shared_ode, res_m, res_t),
               "acb_ode_solve_fuchs_batch")
    return entry_major_to_instance_major(unpack_entries(L, res_m, res_t), batch, n + 1)


def acb_ode_solve_frobenius1_batch(ode: Balls, rho: Balls, order: int, degree: int, n: int, batch: int,
                                   shared_ode: bool = True, shared_rho: bool = False,
                                   valuation: int | None = None, driver=None) -> Balls:
    """Batched ``_acb_ode_solve_frobenius`` with a zero right-hand side and ``M == 1``
    (reference src/frobenius_solver.c:72-121).  ``rho``: ``[batch or 1]`` complex balls."""
    drv = driver or default_driver()
    L = ode.L
    v = acb_ode_valuation(ode, order, degree) if valuation is None else valuation
    ode_m, ode_t = _ode_to_device(ode, order, degree, batch, shared_ode)
    nr = 1 if shared_rho else batch
    rho_m, rho_t = pack_entries(rho, 1, 2 * nr)
    res_m = np.zeros((n + 1, L, 2 * batch), np.uint32)
    res_t = np.zeros((n + 1, 2 * batch, 4), np.int32)
    _lib.check(drv.frobenius1(L, order, degree, v, n, batch, ode_m, ode_t, shared_ode, rho_m, rho_t,
                              shared_rho, res_m, res_t), "acb_ode_solve_frobenius1_batch")
    return entry_major_to_instance_major(unpack_entries(L, res_m, res_t), batch, n + 1)


def acb_poly_evaluate_batch(f: Balls, pts: Balls, nder: int = 1, driver=None) -> Balls:
    """Horner evaluation of one series at many points (``acb_poly_evaluate`` inside
    ``acb_ode_solution_evaluate``, reference src/acb_ode_solution.c:40); with ``nder > 1`` also the
    normalised derivatives ``f^(k)(pt)/k!`` that ``analytic_continuation`` keeps
    (reference src/fuchs_solver.c:159).  Returns ``[npts][nder]`` complex balls."""
    drv = driver or default_driver()
    L = f.L
    length, npts = f.n_slots // 2, pts.n_slots // 2
    f_m, f_t = pack_entries(f, length, 2)
    p_m, p_t = pack_entries(pts, 1, 2 * npts)
    o_m = np.zeros((nder, L, 2 * npts), np.uint32)
    o_t = np.zeros((nder, 2 * npts, 4), np.int32)
    _lib.check(drv.horner(L, length, npts, nder, f_m, f_t, True, p_m, p_t, o_m, o_t), "acb_poly_evaluate_batch")
    return entry_major_to_instance_major(unpack_entries(L, o_m, o_t), npts, nder)


def ew(op: int, x: Balls, y: Balls | None = None, k=None, driver=None) -> Balls:
    """Elementwise ball operation over ``n`` complex balls (acb_add/sub/mul/div/inv/mul_si/add_si)."""
    drv = driver or default_driver()
    L, n = x.L, x.n_slots // 2
    x_m, x_t = pack_entries(x, 1, 2 * n)
    y_m = y_t = None
    if y is not None:
        y_m, y_t = pack_entries(y, 1, 2 * n)
    kk = None
    if k is not None:
        kk = np.ascontiguousarray(np.broadcast_to(np.asarray(k, np.int64), (n,)))
    z_m = np.zeros((1, L, 2 * n), np.uint32)
    z_t = np.zeros((1, 2 * n, 4), np.int32)
    _lib.check(drv.ew(op, L, n, z_m, z_t, x_m, x_t, y_m, y_t, kk), "ew")
    return unpack_entries(L, z_m, z_t)
