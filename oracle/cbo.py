"""ctypes binding of the CPU oracle (oracle/_build/libcbo*.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (cascade_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from cascade_b200.format import Balls

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict = {}

OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_INV = 0, 1, 2, 3, 4
OP_MUL_SI, OP_ADD_SI = 16, 17

_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build() -> None:
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib(plain: bool = False) -> C.CDLL:
    name = "libcbo_plain.so" if plain else "libcbo.so"
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(_HERE, "_build", name)
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.cbo_backend.restype = C.c_char_p
    L.cbo_ew_binary.argtypes = [C.c_int, C.c_int, C.c_int64, _u32p, _i32p, _u32p, _i32p, _u32p, _i32p]
    L.cbo_ew_scalar_si.argtypes = [C.c_int, C.c_int, C.c_int64, _u32p, _i32p, _u32p, _i32p, _i64p]
    L.cbo_fuchs_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int,
                                  _u32p, _i32p, _u32p, _i32p, C.c_int]
    L.cbo_frobenius1_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int,
                                       _u32p, _i32p, _u32p, _i32p, _u32p, _i32p, C.c_int]
    L.cbo_horner_batch.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int, _u32p, _i32p, _u32p,
                                   _i32p, _u32p, _i32p, C.c_int]
    L.cbo_ode_shift_flat.argtypes = [C.c_int, C.c_int, C.c_int, _u32p, _i32p, _u32p, _i32p, _u32p, _i32p]
    L.cbo_ode_solves_flat.argtypes = [C.c_int, C.c_int, C.c_int, _u32p, _i32p, _u32p, _i32p,
                                      C.c_int64, C.c_int64]
    L.cbo_continuation_flat.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, _u32p, _i32p,
                                        _u32p, _i32p, _u32p, _i32p]
    L.cbo_continuation_series_flat.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, _u32p, _i32p,
                                               _u32p, _i32p, _u32p, _i32p, _u32p, _i32p]
    L.cbo_taylor_shift_flat.argtypes = [C.c_int, C.c_int64, C.c_int64, _u32p, _i32p, _u32p, _i32p, C.c_int]
    L.cbo_radius_flat.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, _u32p, _i32p, C.c_int64, _u32p, _i32p, _i32p]
    L.cbo_truncation_flat.argtypes = [C.c_int, C.c_int64, _u32p, _i32p, _u32p, _i32p, C.c_int64, _u32p, _i32p, _i64p]
    L.cbo_example_flat.argtypes = [C.c_int, C.c_int, C.c_uint64, _u32p, _i32p, _u32p, _i32p]
    L.cbo_ode_apply_flat.argtypes = [C.c_int, C.c_int, C.c_int, _u32p, _i32p, _u32p, _i32p, C.c_int64, _u32p, _i32p]
    L.cbo_ew_trans.argtypes = [C.c_int, C.c_int, C.c_int64, _u32p, _i32p, _u32p, _i32p]
    L.cbo_solution_evaluate_flat.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, _u32p, _i32p, _u32p, _i32p, C.c_int,
                                             C.c_uint64, _u32p, _i32p, _u32p, _i32p]
    L.cbo_frobenius1_rhs_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, _u32p, _i32p,
                                           _u32p, _i32p, _u32p, _i32p, C.c_int64, _u32p, _i32p]
    L.cbo_frobenius_general_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, _u32p, _i32p,
                                              _u32p, _i32p, C.c_int, C.c_int64, C.c_int64, _u32p, _i32p, C.c_int]
    L.cbo_indicial_polynomial_flat.argtypes = [C.c_int, C.c_int, C.c_int, _u32p, _i32p, C.c_int64, C.c_int64,
                                               _u32p, _i32p]
    L.cbo_indicial_polynomial_evaluate_flat.argtypes = [C.c_int, C.c_int, C.c_int, _u32p, _i32p, C.c_int64, _u32p,
                                                        _i32p, C.c_int64, _u32p, _i32p]
    L.cbo_solution_op_flat.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                       _u32p, _i32p, _u32p, _i32p, _u32p, _i32p, C.c_int64]
    _LIBS[name] = L
    return L


def backend(plain: bool = False) -> str:
    return lib(plain).cbo_backend().decode()


def ew_binary(op: int, x: Balls, y: Balls | None = None, plain: bool = False) -> Balls:
    y = x if y is None else y
    z = Balls.zeros(x.n_slots, x.L)
    rc = lib(plain).cbo_ew_binary(op, x.L, x.n_slots // 2, z.mant, z.meta, x.mant, x.meta, y.mant, y.meta)
    assert rc == 0
    return z


def ew_scalar_si(op: int, x: Balls, k, plain: bool = False) -> Balls:
    z = Balls.zeros(x.n_slots, x.L)
    k = np.ascontiguousarray(np.broadcast_to(np.asarray(k, np.int64), (x.n_slots // 2,)))
    rc = lib(plain).cbo_ew_scalar_si(op, x.L, x.n_slots // 2, z.mant, z.meta, x.mant, x.meta, k)
    assert rc == 0
    return z


def _series_with_init(init: Balls, batch: int, n: int, ninit: int) -> Balls:
    """[batch][n+1] complex balls with the first ninit of each row taken from init[batch][ninit]."""
    L = init.L
    res = Balls.zeros(batch * (n + 1) * 2, L)
    rm = res.mant.reshape(batch, n + 1, 2, L)
    rt = res.meta.reshape(batch, n + 1, 2, 4)
    rm[:, :ninit] = init.mant.reshape(batch, ninit, 2, L)
    rt[:, :ninit] = init.meta.reshape(batch, ninit, 2, 4)
    return res


def fuchs_batch(order: int, degree: int, n: int, ode: Balls, init: Balls, batch: int,
                shared_ode: bool, nthreads: int = 1, plain: bool = False):
    """init: [batch][ninit] complex balls (ninit = -valuation).  Returns (rc, res[batch][n+1])."""
    ninit = init.n_slots // (2 * batch)
    res = _series_with_init(init, batch, n, ninit)
    rc = lib(plain).cbo_fuchs_batch(ode.L, order, degree, n, batch, int(shared_ode), ode.mant,
                                    ode.meta, res.mant, res.meta, nthreads)
    return rc, res


def frobenius1_batch(order: int, degree: int, n: int, ode: Balls, rho: Balls, batch: int,
                     shared_ode: bool, nthreads: int = 1, plain: bool = False) -> Balls:
    res = Balls.zeros(batch * (n + 1) * 2, ode.L)
    rc = lib(plain).cbo_frobenius1_batch(ode.L, order, degree, n, batch, int(shared_ode), ode.mant,
                                         ode.meta, rho.mant, rho.meta, res.mant, res.meta, nthreads)
    assert rc == 0
    return res


def horner_batch(f: Balls, pts: Balls, nder: int = 1, nthreads: int = 1, plain: bool = False) -> Balls:
    npts = pts.n_slots // 2
    out = Balls.zeros(npts * nder * 2, f.L)
    rc = lib(plain).cbo_horner_batch(f.L, f.n_slots // 2, npts, nder, f.mant, f.meta, pts.mant,
                                     pts.meta, out.mant, out.meta, nthreads)
    assert rc == 0
    return out


def ode_shift(order: int, degree: int, ode: Balls, a: Balls, plain: bool = False) -> Balls:
    out = Balls.zeros(ode.n_slots, ode.L)
    rc = lib(plain).cbo_ode_shift_flat(ode.L, order, degree, ode.mant, ode.meta, a.mant, a.meta,
                                       out.mant, out.meta)
    assert rc == 0
    return out


def ode_solves(order: int, degree: int, ode: Balls, res: Balls, deg: int, plain: bool = False) -> bool:
    return bool(lib(plain).cbo_ode_solves_flat(ode.L, order, degree, ode.mant, ode.meta, res.mant,
                                               res.meta, res.n_slots // 2, deg))


def example(which: str, L: int, n: int = 0, params: Balls | None = None, plain: bool = False) -> Balls:
    """legendre(n) | bessel(nu) | hypgeom(a, b, c): the 3x3 operator as 9 complex balls."""
    code = {"legendre": 0, "bessel": 1, "hypgeom": 2}[which]
    params = Balls.zeros(6, L) if params is None else params.copy()
    out = Balls.zeros(18, L)
    rc = lib(plain).cbo_example_flat(code, L, n, params.mant, params.meta, out.mant, out.meta)
    assert rc == 0
    return out


def continuation(order: int, degree: int, n: int, ode: Balls, path: Balls, y0: Balls, batch: int,
                 plain: bool = False) -> Balls:
    """analytic_continuation for ``batch`` solutions: y0 is [batch][order]; returns the same shape."""
    y = y0.copy()
    rc = lib(plain).cbo_continuation_flat(ode.L, order, degree, n, batch, path.n_slots // 2, ode.mant, ode.meta,
                                          path.mant, path.meta, y.mant, y.meta)
    assert rc == 0
    return y


def continuation_series(order: int, degree: int, n: int, ode: Balls, path: Balls, y0: Balls, batch: int,
                        plain: bool = False):
    """analytic_continuation keeping what the reference leaves in ``res``: returns (rc, y[batch][order],
    series[batch][n+1]) -- the complete series re-expanded at the last corner (acb_poly_taylor_shift, Horner form).
    rc = -2: a corner where the shifted operator has valuation >= 0."""
    y = y0.copy()
    full = Balls.zeros(batch * (n + 1) * 2, ode.L)
    rc = lib(plain).cbo_continuation_series_flat(ode.L, order, degree, n, batch, path.n_slots // 2, ode.mant, ode.meta,
                                                 path.mant, path.meta, y.mant, y.meta, full.mant, full.meta)
    return rc, y, full


def taylor_shift(f: Balls, a: Balls, batch: int, plain: bool = False) -> Balls:
    """acb_poly_taylor_shift (Horner form) of ``batch`` series f[batch][len] by a[batch] (or one shared ball)."""
    out = f.copy()
    length = f.n_slots // (2 * batch)
    rc = lib(plain).cbo_taylor_shift_flat(f.L, length, batch, out.mant, out.meta, a.mant, a.meta,
                                          int(a.n_slots == 2))
    assert rc == 0
    return out


def radius_of_convergence(order: int, degree: int, ode: Balls, batch: int, n: int, shared: bool = False, plain: bool = False):
    """radius_of_convergence (reference src/fuchs_solver.c:3-54) for ``batch`` operators: returns (balls[batch] with a
    zero imaginary part, n_done int32[batch]; n_done = -1: no singularity away from zero, ball indeterminate)."""
    out = Balls.zeros(2 * batch, ode.L)
    nd = np.zeros(batch, np.int32)
    rc = lib(plain).cbo_radius_flat(ode.L, order, degree, batch, int(shared), ode.mant, ode.meta, n, out.mant, out.meta, nd)
    assert rc == 0
    return out, nd


def truncation_order(eta: Balls, alpha: Balls, bits: int, plain: bool = False):
    """truncation_order (reference src/fuchs_solver.c:56-94) for a batch of (eta, alpha): returns (N balls, counts)."""
    batch = eta.n_slots // 2
    nb = Balls.zeros(2 * batch, eta.L)
    out = np.zeros(batch, np.int64)
    rc = lib(plain).cbo_truncation_flat(eta.L, batch, eta.mant, eta.meta, alpha.mant, alpha.meta, bits, nb.mant, nb.meta, out)
    assert rc == 0
    return nb, out


def frobenius1_rhs_batch(order: int, degree: int, n: int, ode: Balls, rho: Balls, rhs: Balls, batch: int,
                         shared_ode: bool, plain: bool = False) -> Balls:
    """_acb_ode_solve_frobenius with right-hand sides rhs[batch][rhs_len] (reference src/frobenius_solver.c:72-121)."""
    res = Balls.zeros(batch * (n + 1) * 2, ode.L)
    rc = lib(plain).cbo_frobenius1_rhs_batch(ode.L, order, degree, n, batch, int(shared_ode), ode.mant, ode.meta,
                                             rho.mant, rho.meta, rhs.mant, rhs.meta, rhs.n_slots // (2 * batch),
                                             res.mant, res.meta)
    assert rc == 0
    return res


def frobenius_general_batch(order: int, degree: int, n: int, ode: Balls, rho: Balls, batch: int, shared_ode: bool,
                            shared_rho: bool, mul: int, M: int, nthreads: int = 1, plain: bool = False) -> Balls:
    """acb_ode_solve_frobenius for M = mul + alpha > 1 (reference src/frobenius_solver.c:131-205).
    Returns sol->gens as [batch][M][n+1] complex balls."""
    gens = Balls.zeros(batch * M * (n + 1) * 2, ode.L)
    rc = lib(plain).cbo_frobenius_general_batch(ode.L, order, degree, n, batch, int(shared_ode), ode.mant, ode.meta,
                                                rho.mant, rho.meta, int(shared_rho), mul, M, gens.mant, gens.meta,
                                                nthreads)
    assert rc == 0
    return gens


def indicial_polynomial(order: int, degree: int, ode: Balls, nu: int, shift: int, plain: bool = False) -> Balls:
    out = Balls.zeros(2 * (order + 1), ode.L)
    rc = lib(plain).cbo_indicial_polynomial_flat(ode.L, order, degree, ode.mant, ode.meta, nu, shift, out.mant, out.meta)
    assert rc == 0
    return out


def indicial_polynomial_evaluate(order: int, degree: int, ode: Balls, nu: int, rho: Balls, shift: int,
                                 plain: bool = False) -> Balls:
    out = Balls.zeros(2, ode.L)
    rc = lib(plain).cbo_indicial_polynomial_evaluate_flat(ode.L, order, degree, ode.mant, ode.meta, nu, rho.mant,
                                                          rho.meta, shift, out.mant, out.meta)
    assert rc == 0
    return out


def solution_op(which: str, rho: Balls, gens: Balls, batch: int, mul: int, M: int, f: Balls | None = None,
                nu: int = 0, plain: bool = False):
    """_acb_ode_solution_extend / _update / _normalize on [batch][M][cap] series; f: [batch][fcap]
    coefficients of the polynomial in rho (consumed).  Returns (gens, f) after the call."""
    code = {"extend": 0, "update": 1, "normalize": 2}[which]
    L = gens.L
    cap = gens.n_slots // (2 * batch * M)
    g = gens.copy()
    ff = Balls.zeros(2 * batch, L) if f is None else f.copy()
    fcap = ff.n_slots // (2 * batch)
    rc = lib(plain).cbo_solution_op_flat(code, L, batch, mul, M, cap, nu, rho.mant, rho.meta, g.mant, g.meta,
                                         ff.mant, ff.meta, fcap)
    assert rc == 0
    return g, ff


def ew_trans(which: str, x: Balls, plain: bool = False) -> Balls:
    """acb_exp / acb_log of n complex balls (the oracle's algorithms, see oracle/cbo_trans.c)."""
    z = Balls.zeros(x.n_slots, x.L)
    rc = lib(plain).cbo_ew_trans({"exp": 0, "log": 1}[which], x.L, x.n_slots // 2, z.mant, z.meta, x.mant, x.meta)
    assert rc == 0
    return z


def solution_evaluate(gens: Balls, M: int, rho: Balls, pow_mode: int, pow_e: int, pts: Balls, plain: bool = False) -> Balls:
    """acb_ode_solution_evaluate (reference src/acb_ode_solution.c:24-58) of one solution (gens[M][cap], rho)
    at every point of pts."""
    cap = gens.n_slots // (2 * M)
    npts = pts.n_slots // 2
    out = Balls.zeros(2 * npts, gens.L)
    rc = lib(plain).cbo_solution_evaluate_flat(gens.L, M, cap, npts, gens.mant, gens.meta, rho.mant, rho.meta, pow_mode,
                                               pow_e, pts.mant, pts.meta, out.mant, out.meta)
    assert rc == 0
    return out


def ode_apply(order: int, degree: int, ode: Balls, res: Balls, plain: bool = False) -> Balls:
    """acb_ode_apply (reference src/acb_ode.c:185-206): the residual series, len(res) + degree coefficients."""
    length = res.n_slots // 2
    out = Balls.zeros(2 * (length + degree), ode.L)
    rc = lib(plain).cbo_ode_apply_flat(ode.L, order, degree, ode.mant, ode.meta, res.mant, res.meta, length, out.mant, out.meta)
    assert rc == 0
    return out
