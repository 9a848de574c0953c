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

    def fuchs_intop(self, L, order, degree, valuation, n, batch, ode_int, shared, res_m, res_t):
        return _lib.lib().cb200_fuchs_intop_batch_host(L, order, degree, valuation, n, batch, ode_int, int(shared),
                                                       res_m, res_t, self.device)

    def continuation(self, L, order, degree, n, batch, path_len, ode_m, ode_t, path_m, path_t, y_m, y_t):
        return _lib.lib().cb200_continuation_host(L, order, degree, n, batch, path_len, ode_m, ode_t, path_m, path_t,
                                                  y_m, y_t, self.device)

    def continuation_series(self, L, order, degree, n, batch, path_len, ode_m, ode_t, path_m, path_t, y_m, y_t, r_m, r_t):
        return _lib.lib().cb200_continuation_series_host(L, order, degree, n, batch, path_len, ode_m, ode_t, path_m, path_t,
                                                         y_m, y_t, r_m, r_t, self.device)

    def ode_shift(self, L, order, degree, batch, ode_m, ode_t, shared, a_m, a_t, shared_a, o_m, o_t):
        return _lib.lib().cb200_ode_shift_host(L, order, degree, batch, ode_m, ode_t, int(shared), a_m, a_t, int(shared_a),
                                               o_m, o_t, self.device)

    def taylor_shift(self, L, length, batch, f_m, f_t, a_m, a_t, shared_a):
        return _lib.lib().cb200_taylor_shift_host(L, length, batch, f_m, f_t, a_m, a_t, int(shared_a), self.device)

    def radius(self, L, order, degree, batch, ode_m, ode_t, shared, n, o_m, o_t, n_done):
        return _lib.lib().cb200_radius_of_convergence_host(L, order, degree, batch, ode_m, ode_t, int(shared), n, o_m, o_t,
                                                           n_done, self.device)

    def truncation(self, L, batch, e_m, e_t, a_m, a_t, bits, n_m, n_t, n_out):
        return _lib.lib().cb200_truncation_order_host(L, batch, e_m, e_t, a_m, a_t, bits, n_m, n_t, n_out, self.device)

    def frobenius1(self, L, order, degree, valuation, n, batch, ode_m, ode_t, shared, rho_m, rho_t,
                   shared_rho, res_m, res_t, rhs_m=None, rhs_t=None, rhs_len=0, shared_rhs=False):
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        return _lib.lib().cb200_frobenius1_rhs_batch_host(L, order, degree, valuation, n, batch, ode_m, ode_t,
                                                          int(shared), rho_m, rho_t, int(shared_rho), vp(rhs_m),
                                                          vp(rhs_t), rhs_len, int(shared_rhs), res_m, res_t,
                                                          self.device)

    def frobenius_general(self, L, order, degree, valuation, n, batch, ode_m, ode_t, shared, rho_m, rho_t,
                          shared_rho, mul, M, gens_m, gens_t):
        return _lib.lib().cb200_frobenius_general_batch_host(L, order, degree, valuation, n, batch, ode_m, ode_t,
                                                             int(shared), rho_m, rho_t, int(shared_rho), mul, M,
                                                             gens_m, gens_t, self.device)

    def solution_op(self, which, L, batch, mul, M, cap, nu, rho_m, rho_t, gens_m, gens_t, f_m, f_t, fcap):
        return _lib.lib().cb200_solution_op_host(which, L, batch, mul, M, cap, nu, rho_m, rho_t, gens_m, gens_t,
                                                 f_m, f_t, fcap, self.device)

    def indicial_polynomial(self, L, order, degree, valuation, batch, ode_m, ode_t, shared, nu, shift, out_m, out_t):
        return _lib.lib().cb200_indicial_polynomial_host(L, order, degree, valuation, batch, ode_m, ode_t, int(shared),
                                                         nu, shift, out_m, out_t, self.device)

    def solution_evaluate(self, L, M, cap, npts, g_m, g_t, rho_m, rho_t, pow_mode, pow_e, p_m, p_t, o_m, o_t):
        return _lib.lib().cb200_solution_evaluate_host(L, M, cap, npts, g_m, g_t, rho_m, rho_t, pow_mode, pow_e, p_m, p_t,
                                                       o_m, o_t, self.device)

    def trans(self, which, L, n, z_m, z_t, x_m, x_t):
        return _lib.lib().cb200_trans_host(which, L, n, z_m, z_t, x_m, x_t, self.device)

    def ode_apply(self, L, order, degree, batch, ode_m, ode_t, shared, in_m, in_t, len_in, o_m, o_t, out_len, deg, solved):
        return _lib.lib().cb200_ode_apply_host(L, order, degree, batch, ode_m, ode_t, int(shared), in_m, in_t, len_in, o_m,
                                               o_t, out_len, deg, solved, self.device)

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


def acb_ode_valuations(ode: Balls, order: int, degree: int) -> np.ndarray:
    """``acb_ode_valuation`` of every operator of an ``[instances][entries]`` array at once (one vectorised
    exact-zero scan of the ball records): int64 ``[instances]``."""
    ne = (order + 1) * (degree + 1)
    meta = ode.meta.reshape(-1, ne, 2, 4)
    exact_zero = ((meta[..., 1] & FLAG_ZERO) != 0) & (meta[..., 2] == 0)
    zero = exact_zero.all(axis=-1).reshape(-1, order + 1, degree + 1)
    # ord_z p_i = index of the first coefficient that is not an exact zero (degree + 1 for the zero row)
    nz = ~zero
    first = np.where(nz.any(axis=-1), nz.argmax(axis=-1), degree + 1)
    val = (first - np.arange(order + 1)[None, :]).min(axis=-1)
    return np.minimum(val, degree).astype(np.int64)


def batch_valuation(ode: Balls, order: int, degree: int, shared: bool, what: str) -> int:
    """The ONE valuation a batched launch runs with.  The kernels take a single scalar valuation per launch
    (include/cascade_b200.h), so operators that differ between instances must all have the same valuation;
    exact-zero coefficients (``acb_ode_random`` emits them) can make them differ, and such a batch is rejected
    here instead of being solved with the wrong recurrence window.  Split it by valuation and call per group."""
    if shared:
        return acb_ode_valuation(ode, order, degree)
    vals = acb_ode_valuations(ode, order, degree)
    if vals.size and (vals != vals[0]).any():
        uniq = np.unique(vals)
        raise _lib.CascadeB200Error(
            f"{what}: the operators of this batch have different valuations {uniq.tolist()} "
            f"(instance {int(np.argmax(vals != vals[0]))} differs from instance 0); a batched launch runs with one "
            "valuation -- group the instances by acb_ode_valuations() and solve each group")
    return int(vals[0]) if vals.size else acb_ode_valuation(ode, order, degree)


def _ode_to_device(ode: Balls, order: int, degree: int, batch: int, shared: bool):
    ne = (order + 1) * (degree + 1)
    nb = 1 if shared else batch
    assert ode.n_slots == nb * ne * 2, "operator array has the wrong size"
    return pack_entries(instance_major_to_entry_major(ode, nb, ne), ne, 2 * nb)


def integer_operator(ode: Balls, order: int, degree: int, n: int):
    """If every coefficient of every operator is an exact real integer and the integer
    multipliers of the recurrence stay below 2^62, return them as int64 ``[instances][entries]``;
    else None.  (Host-side inspection of the ball records only.)"""
    ne = (order + 1) * (degree + 1)
    L = ode.L
    meta = ode.meta.reshape(-1, ne, 2, 4)
    mant = ode.mant.reshape(-1, ne, 2, L)
    flags = meta[..., 1]
    if (meta[..., 2] != 0).any() or (flags & FLAG_NAN).any():
        return None  # a radius somewhere
    if not (((flags[:, :, 1] & FLAG_ZERO) != 0)).all():
        return None  # an imaginary part
    re_flags, re_exp, re_m = flags[:, :, 0], meta[:, :, 0, 0].astype(np.int64), mant[:, :, 0]
    nz = (re_flags & FLAG_ZERO) == 0
    if (re_m[..., : L - 2] != 0).any():
        return None  # more than 64 significant bits
    top = (re_m[..., L - 1].astype(np.uint64) << np.uint64(32)) | re_m[..., L - 2].astype(np.uint64)
    if (nz & ((re_exp < 1) | (re_exp > 62))).any():
        return None
    sh = np.where(nz, 64 - re_exp, 0).astype(np.uint64)
    val = np.where(nz, top >> sh, 0).astype(np.uint64)
    if (np.where(nz, (val << sh) != top, False)).any():
        return None  # fractional bits
    out = val.astype(np.int64)
    out = np.where((re_flags & 1) != 0, -out, out)
    bound = float(np.abs(out).max(initial=0)) * float(n + 1) ** order * (order + 1)
    if bound >= 2.0 ** 62:
        return None
    return np.ascontiguousarray(out)


# ---------------------------------------------------------------- solvers
def acb_ode_solve_fuchs_batch(ode: Balls, init: Balls, order: int, degree: int, n: int, batch: int,
                              shared_ode: bool = True, valuation: int | None = None, driver=None,
                              use_intop: bool = True) -> Balls:
    """Batched ``acb_ode_solve_fuchs(res, ODE, n, 32*L)`` (reference src/fuchs_solver.c:96-145).

    ``init``: ``[batch][-valuation]`` complex balls (``res[0..-v-1]`` of each call).
    Returns ``[batch][n+1]`` complex balls.  Raises for ``valuation >= 0`` (undefined in the
    reference)."""
    drv = driver or default_driver()
    L = ode.L
    v = batch_valuation(ode, order, degree, shared_ode, "acb_ode_solve_fuchs_batch") if valuation is None else valuation
    if v >= 0:
        _lib.check(_lib.EVALUATION, "acb_ode_solve_fuchs_batch")
    ninit = -v
    assert init.n_slots == batch * ninit * 2, "need -valuation initial values per instance"
    assert n >= ninit - 1
    res = Balls.zeros(batch * (n + 1) * 2, L)
    res.mant.reshape(batch, n + 1, 2, L)[:, :ninit] = init.mant.reshape(batch, ninit, 2, L)
    res.meta.reshape(batch, n + 1, 2, 4)[:, :ninit] = init.meta.reshape(batch, ninit, 2, 4)
    res_m, res_t = pack_entries(instance_major_to_entry_major(res, batch, n + 1), n + 1, 2 * batch)
    ints = integer_operator(ode, order, degree, n) if (use_intop and not shared_ode) else None
    if ints is not None and (order + 1) * (degree + 1) <= 36:
        # integer operators (Legendre sweeps): O(L) kernel, same balls
        _lib.check(drv.fuchs_intop(L, order, degree, v, n, batch, np.ascontiguousarray(ints.T), False, res_m, res_t),
                   "acb_ode_solve_fuchs_batch (integer operators)")
    else:
        ode_m, ode_t = _ode_to_device(ode, order, degree, batch, shared_ode)
        _lib.check(drv.fuchs(L, order, degree, v, n, batch, ode_m, ode_t, shared_ode, res_m, res_t),
                   "acb_ode_solve_fuchs_batch")
    return entry_major_to_instance_major(unpack_entries(L, res_m, res_t), batch, n + 1)


def acb_ode_solve_frobenius1_batch(ode: Balls, rho: Balls, order: int, degree: int, n: int, batch: int,
                                   shared_ode: bool = True, shared_rho: bool = False,
                                   valuation: int | None = None, driver=None, rhs: Balls | None = None) -> Balls:
    """Batched ``_acb_ode_solve_frobenius`` with ``M == 1`` (reference src/frobenius_solver.c:72-121).
    ``rho``: ``[batch or 1]`` complex balls.  ``rhs``: optional right-hand side series, ``[batch][rhs_len]``;
    instances whose rhs is the zero polynomial take the homogeneous branch (g_0 = 1), the others use
    rho - valuation and g_0 = rhs_0 / f_0(rho) as the reference does (:86-96)."""
    drv = driver or default_driver()
    L = ode.L
    v = batch_valuation(ode, order, degree, shared_ode, "acb_ode_solve_frobenius1_batch") if valuation is None else valuation
    ode_m, ode_t = _ode_to_device(ode, order, degree, batch, shared_ode)
    nr = 1 if shared_rho else batch
    rho_m, rho_t = pack_entries(rho, 1, 2 * nr)
    res_m = np.zeros((n + 1, L, 2 * batch), np.uint32)
    res_t = np.zeros((n + 1, 2 * batch, 4), np.int32)
    if rhs is None:
        rc = drv.frobenius1(L, order, degree, v, n, batch, ode_m, ode_t, shared_ode, rho_m, rho_t, shared_rho, res_m, res_t)
    else:
        hl = rhs.n_slots // (2 * batch)
        h_m, h_t = pack_entries(instance_major_to_entry_major(rhs, batch, hl), hl, 2 * batch)
        rc = drv.frobenius1(L, order, degree, v, n, batch, ode_m, ode_t, shared_ode, rho_m, rho_t, shared_rho, res_m,
                            res_t, h_m, h_t, hl, False)
    _lib.check(rc, "acb_ode_solve_frobenius1_batch")
    return entry_major_to_instance_major(unpack_entries(L, res_m, res_t), batch, n + 1)


def acb_ode_solve_frobenius_batch(ode: Balls, rho: Balls, mul: int, alpha: int, order: int, degree: int, n: int,
                                  batch: int, shared_ode: bool = True, shared_rho: bool = False,
                                  valuation: int | None = None, driver=None) -> Balls:
    """Batched ``acb_ode_solve_frobenius(sol, ODE, n, prec)`` after ``acb_ode_solution_init(sol, rho, mul,
    alpha)`` (reference src/frobenius_solver.c:123-205).  Returns ``sol->gens`` as ``[batch][M][n+1]``
    complex balls, ``M = mul + alpha``: the M = 1 recurrence or the polynomial-in-rho (logarithmic) one."""
    drv = driver or default_driver()
    M = mul + alpha
    if M == 1:
        return acb_ode_solve_frobenius1_batch(ode, rho, order, degree, n, batch, shared_ode, shared_rho, valuation, drv)
    L = ode.L
    v = batch_valuation(ode, order, degree, shared_ode, "acb_ode_solve_frobenius_batch") if valuation is None else valuation
    ode_m, ode_t = _ode_to_device(ode, order, degree, batch, shared_ode)
    rho_m, rho_t = pack_entries(rho, 1, 2 * (1 if shared_rho else batch))
    ng = M * (n + 1)
    g_m = np.zeros((ng, L, 2 * batch), np.uint32)
    g_t = np.zeros((ng, 2 * batch, 4), np.int32)
    _lib.check(drv.frobenius_general(L, order, degree, v, n, batch, ode_m, ode_t, shared_ode, rho_m, rho_t,
                                     shared_rho, mul, M, g_m, g_t), "acb_ode_solve_frobenius_batch")
    return entry_major_to_instance_major(unpack_entries(L, g_m, g_t), batch, ng)


SOL_EXTEND, SOL_UPDATE, SOL_NORMALIZE = 0, 1, 2


def acb_ode_solution_op_batch(which: int, rho: Balls, gens: Balls, batch: int, mul: int, M: int,
                              f: Balls | None = None, nu: int = 0, driver=None):
    """``_acb_ode_solution_extend`` / ``_update`` / ``_normalize`` (reference src/acb_ode_solution.c:60-123)
    for a batch of solutions.  ``gens``: ``[batch][M][cap]``; ``f``: ``[batch][fcap]`` coefficients of the
    polynomial in rho, consumed as in the reference.  Returns ``(gens, f)`` after the call."""
    drv = driver or default_driver()
    L = gens.L
    cap = gens.n_slots // (2 * batch * max(M, 1)) if M > 0 else 0
    ff = Balls.zeros(2 * batch, L) if f is None else f.copy()  # the call consumes its polynomial
    fcap = ff.n_slots // (2 * batch)
    g_m, g_t = pack_entries(instance_major_to_entry_major(gens.copy(), batch, M * cap), M * cap, 2 * batch)
    f_m, f_t = pack_entries(instance_major_to_entry_major(ff, batch, fcap), fcap, 2 * batch)
    rho_m, rho_t = pack_entries(rho, 1, 2 * batch)
    _lib.check(drv.solution_op(which, L, batch, mul, M, cap, nu, rho_m, rho_t, g_m, g_t, f_m, f_t, fcap),
               "acb_ode_solution_op_batch")
    return (entry_major_to_instance_major(unpack_entries(L, g_m, g_t), batch, M * cap),
            entry_major_to_instance_major(unpack_entries(L, f_m, f_t), batch, fcap))


def indicial_polynomial_batch(ode: Balls, order: int, degree: int, nu: int, shift: int, batch: int,
                              shared_ode: bool = True, valuation: int | None = None, driver=None) -> Balls:
    """``indicial_polynomial(result, ODE, nu, shift, prec)`` (reference src/frobenius_solver.c:4-39):
    ``[batch][order+1]`` coefficients of f_nu(rho + shift) as a polynomial in rho."""
    drv = driver or default_driver()
    L = ode.L
    v = batch_valuation(ode, order, degree, shared_ode, "indicial_polynomial_batch") if valuation is None else valuation
    ode_m, ode_t = _ode_to_device(ode, order, degree, batch, shared_ode)
    o_m = np.zeros((order + 2, L, 2 * batch), np.uint32)
    o_t = np.zeros((order + 2, 2 * batch, 4), np.int32)
    _lib.check(drv.indicial_polynomial(L, order, degree, v, batch, ode_m, ode_t, shared_ode, nu, shift, o_m, o_t),
               "indicial_polynomial_batch")
    full = entry_major_to_instance_major(unpack_entries(L, o_m, o_t), batch, order + 2)
    keep = np.arange(batch * (order + 2) * 2).reshape(batch, order + 2, 2)[:, : order + 1].reshape(-1)
    return Balls(L, full.mant[keep].copy(), full.meta[keep].copy())


def analytic_continuation_batch(ode: Balls, path: Balls, y0: Balls, order: int, degree: int, n: int, batch: int,
                                driver=None) -> Balls:
    """``analytic_continuation(res, ODE, path, len, n, bits)`` (reference src/fuchs_solver.c:147-163)
    for ``batch`` solutions of one operator along one path, device resident.  ``y0``: ``[batch][order]``
    Taylor coefficients at ``path[0]``; returns those at the last corner.  Every corner must be an
    ordinary point of the operator."""
    drv = driver or default_driver()
    L = ode.L
    plen = path.n_slots // 2
    ode_m, ode_t = _ode_to_device(ode, order, degree, batch, True)
    path_m, path_t = pack_entries(path, plen, 2)
    y_m, y_t = pack_entries(instance_major_to_entry_major(y0, batch, order), order, 2 * batch)
    _lib.check(drv.continuation(L, order, degree, n, batch, plen, ode_m, ode_t, path_m, path_t, y_m, y_t),
               "analytic_continuation_batch")
    return entry_major_to_instance_major(unpack_entries(L, y_m, y_t), batch, order)


def analytic_continuation_series_batch(ode: Balls, path: Balls, y0: Balls, order: int, degree: int, n: int, batch: int,
                                       driver=None):
    """``analytic_continuation`` keeping everything the reference leaves in ``res`` (src/fuchs_solver.c:159): returns
    (y ``[batch][order]``, series ``[batch][n+1]``), the series being the complete re-expansion at the last corner
    (``acb_poly_taylor_shift``, Horner form)."""
    drv = driver or default_driver()
    L = ode.L
    plen = path.n_slots // 2
    ode_m, ode_t = _ode_to_device(ode, order, degree, batch, True)
    path_m, path_t = pack_entries(path, plen, 2)
    y_m, y_t = pack_entries(instance_major_to_entry_major(y0, batch, order), order, 2 * batch)
    r_m = np.zeros((n + 1, L, 2 * batch), np.uint32)
    r_t = np.zeros((n + 1, 2 * batch, 4), np.int32)
    _lib.check(drv.continuation_series(L, order, degree, n, batch, plen, ode_m, ode_t, path_m, path_t, y_m, y_t, r_m, r_t),
               "analytic_continuation_series_batch")
    return (entry_major_to_instance_major(unpack_entries(L, y_m, y_t), batch, order),
            entry_major_to_instance_major(unpack_entries(L, r_m, r_t), batch, n + 1))


def acb_poly_taylor_shift_batch(f: Balls, a: Balls, batch: int, driver=None) -> Balls:
    """``acb_poly_taylor_shift(f, f, a)`` (reference src/fuchs_solver.c:159) for ``batch`` series ``f[batch][len]``;
    ``a``: ``[batch]`` balls or one shared ball.  Horner form, O(len^2) ball operations per series."""
    drv = driver or default_driver()
    L = f.L
    length = f.n_slots // (2 * batch)
    shared_a = a.n_slots == 2
    f_m, f_t = pack_entries(instance_major_to_entry_major(f, batch, length), length, 2 * batch)
    a_m, a_t = pack_entries(a, 1, 2 if shared_a else 2 * batch)
    _lib.check(drv.taylor_shift(L, length, batch, f_m, f_t, a_m, a_t, shared_a), "acb_poly_taylor_shift_batch")
    return entry_major_to_instance_major(unpack_entries(L, f_m, f_t), batch, length)


def radius_of_convergence_batch(ode: Balls, order: int, degree: int, batch: int, n: int = 20, shared_ode: bool = False,
                                driver=None):
    """``radius_of_convergence(rad, ODE, n, bits)`` (reference src/fuchs_solver.c:3-54) for a batch of operators, rigorous
    on the device.  Returns (balls ``[batch]`` -- real, the imaginary parts exact zeros --, ``n_done`` int32 ``[batch]``:
    Graeffe iterations performed, -1 where the operator has no singularity away from zero: ball indeterminate)."""
    drv = driver or default_driver()
    L = ode.L
    ode_m, ode_t = _ode_to_device(ode, order, degree, batch, shared_ode)
    o_m = np.zeros((1, L, 2 * batch), np.uint32)
    o_t = np.zeros((1, 2 * batch, 4), np.int32)
    nd = np.zeros(batch, np.int32)
    _lib.check(drv.radius(L, order, degree, batch, ode_m, ode_t, shared_ode, n, o_m, o_t, nd), "radius_of_convergence_batch")
    return unpack_entries(L, o_m, o_t), nd


def truncation_order_batch(eta: Balls, alpha: Balls, bits: int, driver=None):
    """``truncation_order(eta, alpha, bits)`` (reference src/fuchs_solver.c:56-94) for a batch of pairs: returns
    (the balls N, the term counts int64 ``[batch]``)."""
    drv = driver or default_driver()
    L, batch = eta.L, eta.n_slots // 2
    e_m, e_t = pack_entries(eta, 1, 2 * batch)
    a_m, a_t = pack_entries(alpha, 1, 2 * batch)
    n_m = np.zeros((1, L, 2 * batch), np.uint32)
    n_t = np.zeros((1, 2 * batch, 4), np.int32)
    out = np.zeros(batch, np.int64)
    _lib.check(drv.truncation(L, batch, e_m, e_t, a_m, a_t, bits, n_m, n_t, out), "truncation_order_batch")
    return unpack_entries(L, n_m, n_t), out


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


def acb_ode_apply_batch(ode: Balls, series: Balls, order: int, degree: int, batch: int, shared_ode: bool = True,
                        deg: int | None = None, driver=None):
    """``acb_ode_apply`` / ``acb_ode_solves`` (reference src/acb_ode.c:185-240) for a batch, on the device.
    ``series``: ``[batch][len]``.  Returns (residual ``[batch][len + degree]``, solved ``bool[batch]``) where solved
    is the reference's acceptance test on the first ``deg`` residual coefficients."""
    drv = driver or default_driver()
    L = ode.L
    length = series.n_slots // (2 * batch)
    out_len = length + degree
    deg = length - order if deg is None else deg
    ode_m, ode_t = _ode_to_device(ode, order, degree, batch, shared_ode)
    in_m, in_t = pack_entries(instance_major_to_entry_major(series, batch, length), length, 2 * batch)
    o_m = np.zeros((out_len, L, 2 * batch), np.uint32)
    o_t = np.zeros((out_len, 2 * batch, 4), np.int32)
    solved = np.zeros(batch, np.int32)
    _lib.check(drv.ode_apply(L, order, degree, batch, ode_m, ode_t, shared_ode, in_m, in_t, length, o_m, o_t, out_len,
                             max(deg, 0), solved), "acb_ode_apply_batch")
    return entry_major_to_instance_major(unpack_entries(L, o_m, o_t), batch, out_len), solved.astype(bool)


def pow_mode_of(rho: Balls):
    """How ``acb_pow(a, rho)`` is computed for this exponent (host-side inspection of one complex ball):
    (0, 0) rho is the exact zero; (1, e) rho is the exact integer 0 <= e < 2^31; (2, 0) general."""
    L = rho.L
    re_t, im_t = rho.meta[0], rho.meta[1]
    exact = re_t[2] == 0 and im_t[2] == 0 and not ((re_t[1] | im_t[1]) & FLAG_NAN)
    if not exact or not (im_t[1] & FLAG_ZERO):
        return 2, 0
    if re_t[1] & FLAG_ZERO:
        return 0, 0
    if (re_t[1] & 1) or not (1 <= re_t[0] <= 31) or rho.mant[0, : L - 1].any():
        return 2, 0
    top = int(rho.mant[0, L - 1])
    e = top >> (32 - int(re_t[0]))
    if (e << (32 - int(re_t[0]))) != top:
        return 2, 0
    return 1, e


def acb_ode_solution_evaluate_batch(gens: Balls, M: int, rho: Balls, pts: Balls, driver=None) -> Balls:
    """``acb_ode_solution_evaluate(out, sol, a, prec)`` (reference src/acb_ode_solution.c:24-58) of one solution
    (``gens``: ``[M][cap]`` series, ``rho``) at every point of ``pts``; ``acb_log`` / ``acb_pow`` run on the device."""
    drv = driver or default_driver()
    L = gens.L
    cap, npts = gens.n_slots // (2 * M), pts.n_slots // 2
    mode, e = pow_mode_of(rho)
    g_m, g_t = pack_entries(gens, M * cap, 2)
    r_m, r_t = pack_entries(rho, 1, 2)
    p_m, p_t = pack_entries(pts, 1, 2 * npts)
    o_m = np.zeros((1, L, 2 * npts), np.uint32)
    o_t = np.zeros((1, 2 * npts, 4), np.int32)
    _lib.check(drv.solution_evaluate(L, M, cap, npts, g_m, g_t, r_m, r_t, mode, e, p_m, p_t, o_m, o_t),
               "acb_ode_solution_evaluate_batch")
    return unpack_entries(L, o_m, o_t)


def acb_exp_batch(x: Balls, driver=None) -> Balls:
    return _trans(0, x, driver)


def acb_log_batch(x: Balls, driver=None) -> Balls:
    return _trans(1, x, driver)


def _trans(which: int, x: Balls, driver=None) -> Balls:
    drv = driver or default_driver()
    L, n = x.L, x.n_slots // 2
    x_m, x_t = pack_entries(x, 1, 2 * n)
    z_m = np.zeros((1, L, 2 * n), np.uint32)
    z_t = np.zeros((1, 2 * n, 4), np.int32)
    _lib.check(drv.trans(which, L, n, z_m, z_t, x_m, x_t), "acb_exp/acb_log")
    return unpack_entries(L, z_m, z_t)


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


# ---------------------------------------------------------------- example operators (src/examples.c)
def _neg(b: Balls) -> Balls:
    out = b.copy()
    nz = (out.meta[:, 1] & (FLAG_ZERO | FLAG_NAN)) == 0
    out.meta[nz, 1] ^= 1
    return out


def _set_entry(ode: Balls, batch: int, ne: int, entry: int, val: Balls) -> None:
    L = ode.L
    ode.mant.reshape(batch, ne, 2, L)[:, entry] = val.mant.reshape(batch, 2, L)
    ode.meta.reshape(batch, ne, 2, 4)[:, entry] = val.meta.reshape(batch, 2, 4)


def _const(values, L: int) -> Balls:
    from .format import from_complex
    return from_complex(list(values), L)


def acb_ode_legendre_batch(ns, L: int) -> Balls:
    """``acb_ode_legendre(ODE, n)`` for every n in ``ns`` (reference src/examples.c:3-11):
    (1 - z^2) y'' - 2 z y' + n(n+1) y.  Exact small integers, no arithmetic needed."""
    ns = [int(n) for n in ns]
    batch = len(ns)
    ode = Balls.zeros(batch * 9 * 2, L)
    _set_entry(ode, batch, 9, 0 * 3 + 0, _const([n * (n + 1) for n in ns], L))
    _set_entry(ode, batch, 9, 1 * 3 + 1, _const([-2] * batch, L))
    _set_entry(ode, batch, 9, 2 * 3 + 0, _const([1] * batch, L))
    _set_entry(ode, batch, 9, 2 * 3 + 2, _const([-1] * batch, L))
    return ode


def acb_ode_bessel_batch(nu: Balls, driver=None) -> tuple[Balls, Balls]:
    """``acb_ode_bessel(ODE, nu, bits)`` for a batch of nu (reference src/examples.c:13-22).
    Returns (ode, nu_squared): the reference overwrites its ``nu`` argument with nu^2 (:20)."""
    L, batch = nu.L, nu.n_slots // 2
    nu2 = ew(OP_MUL, nu, nu, driver=driver)
    ode = Balls.zeros(batch * 9 * 2, L)
    one = _const([1] * batch, L)
    _set_entry(ode, batch, 9, 2 * 3 + 2, one)
    _set_entry(ode, batch, 9, 1 * 3 + 1, one)
    _set_entry(ode, batch, 9, 0 * 3 + 2, one)
    _set_entry(ode, batch, 9, 0 * 3 + 0, _neg(nu2))
    return ode, nu2


def acb_ode_hypgeom_batch(a: Balls, b: Balls, c: Balls, driver=None) -> Balls:
    """``acb_ode_hypgeom(ODE, a, b, c, bits)`` for a batch (reference src/examples.c:24-44):
    z(1-z) y'' + (c - (a+b+1) z) y' - a b y."""
    L, batch = a.L, a.n_slots // 2
    ode = Balls.zeros(batch * 9 * 2, L)
    _set_entry(ode, batch, 9, 2 * 3 + 1, _const([1] * batch, L))
    _set_entry(ode, batch, 9, 2 * 3 + 2, _const([-1] * batch, L))
    _set_entry(ode, batch, 9, 1 * 3 + 0, c)
    t = ew(OP_ADD, _const([1] * batch, L), a, driver=driver)
    t = ew(OP_ADD, t, b, driver=driver)
    _set_entry(ode, batch, 9, 1 * 3 + 1, _neg(t))
    _set_entry(ode, batch, 9, 0 * 3 + 0, _neg(ew(OP_MUL, a, b, driver=driver)))
    return ode


def acb_ode_shift_batch(ode: Balls, a: Balls, order: int, degree: int, driver=None) -> Balls:
    """``acb_ode_shift(out, in, a, bits)`` for a batch (reference src/acb_ode.c:124-135): every row
    p_i(z) becomes p_i(z + a) by the Horner Taylor shift  p[j] += p[j+1] * a, one launch for the batch.
    ``ode``: ``[batch][entries]`` or one operator (then it is moved to each of the ``batch`` points of ``a``)."""
    drv = driver or default_driver()
    L, batch = ode.L, a.n_slots // 2
    ne = (order + 1) * (degree + 1)
    shared = ode.n_slots == 2 * ne and batch != 1
    ode_m, ode_t = _ode_to_device(ode, order, degree, batch, shared)
    a_m, a_t = pack_entries(a, 1, 2 * batch)
    o_m = np.zeros((ne, L, 2 * batch), np.uint32)
    o_t = np.zeros((ne, 2 * batch, 4), np.int32)
    _lib.check(drv.ode_shift(L, order, degree, batch, ode_m, ode_t, shared, a_m, a_t, False, o_m, o_t), "acb_ode_shift_batch")
    return entry_major_to_instance_major(unpack_entries(L, o_m, o_t), batch, ne)
