"""TEST INFRASTRUCTURE: drives tests/host_sim/libcbsim.so (the device thread bodies compiled for
the host) through the same driver interface as cascade_b200.api.HostBufferDriver."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcbsim.so")
_SRC = os.path.join(_HERE, "host_sim.cpp")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_vp = C.c_void_p
_lib = None


def build(force: bool = False) -> None:
    deps = [_SRC] + [os.path.join(_HERE, "..", "..", "cascade_b200", "csrc", f)
                     for f in ("cb_arith.cuh", "cb_solvers.cuh", "cb_mul.cuh", "cb_rr.cuh", "cb_trans.cuh", "cb_intop.cuh", "cb_coop.cuh")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-ffp-contract=off",
                    "-o", _SO, _SRC], check=True)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.cbsim_path_count.argtypes = [C.c_int]
        L.cbsim_path_count.restype = C.c_longlong
        L.cbsim_fuchs_batch.argtypes = [C.c_int] * 4 + [C.c_int64] * 2 + [_u32p, _i32p, C.c_int, _u32p, _i32p]
        L.cbsim_fuchs_batch_range.argtypes = [C.c_int] * 4 + [C.c_int64] * 3 + [_u32p, _i32p, C.c_int, _u32p, _i32p]
        L.cbsim_fuchs_intop_batch.argtypes = [C.c_int] * 4 + [C.c_int64] * 2 + [np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS"), C.c_int, _u32p, _i32p]
        L.cbsim_continuation.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, _u32p, _i32p, _u32p, _i32p, _u32p, _i32p]
        L.cbsim_continuation_series.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, _u32p, _i32p, _u32p, _i32p,
                                                _u32p, _i32p, _u32p, _i32p]
        L.cbsim_ode_shift.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _u32p, _i32p, C.c_int, _u32p, _i32p, C.c_int, _u32p, _i32p]
        L.cbsim_taylor_shift.argtypes = [C.c_int, C.c_int64, C.c_int64, _u32p, _i32p, _u32p, _i32p, C.c_int]
        L.cbsim_radius.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _u32p, _i32p, C.c_int, C.c_int64, _u32p, _i32p, _i32p]
        L.cbsim_truncation.argtypes = [C.c_int, C.c_int64, _u32p, _i32p, _u32p, _i32p, C.c_int64, _u32p, _i32p,
                                       np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")]
        L.cbsim_frobenius1_coop.argtypes = [C.c_int] * 4 + [C.c_int64] * 2 + [_u32p, _i32p, C.c_int, _u32p, _i32p, C.c_int, _u32p, _i32p]
        L.cbsim_horner_coop.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int, _u32p, _i32p, C.c_int, _u32p, _i32p, _u32p, _i32p]
        L.cbsim_frobenius1_batch.argtypes = [C.c_int] * 4 + [C.c_int64] * 2 + [_u32p, _i32p, C.c_int, _u32p, _i32p, C.c_int, _vp, _vp, C.c_int64, C.c_int, _u32p, _i32p]
        L.cbsim_horner_batch.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int, _u32p, _i32p, C.c_int, _u32p, _i32p, _u32p, _i32p]
        L.cbsim_ew.argtypes = [C.c_int, C.c_int, C.c_int64, _u32p, _i32p, _u32p, _i32p, _vp, _vp, _vp]
        L.cbsim_frobenius_general_batch.argtypes = ([C.c_int] * 4 + [C.c_int64] * 2 + [_u32p, _i32p, C.c_int, _u32p, _i32p,
                                                    C.c_int, C.c_int, C.c_int, _u32p, _i32p])
        L.cbsim_solution_op.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64, _u32p, _i32p,
                                        _u32p, _i32p, _u32p, _i32p, C.c_int64]
        L.cbsim_indicial_polynomial.argtypes = [C.c_int] * 4 + [C.c_int64, _u32p, _i32p, C.c_int, C.c_int64, C.c_int64,
                                                                _u32p, _i32p]
        L.cbsim_solution_evaluate.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, _u32p, _i32p, _u32p, _i32p, C.c_int,
                                              C.c_uint64, _u32p, _i32p, _u32p, _i32p]
        L.cbsim_ode_apply.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _u32p, _i32p, C.c_int, _u32p, _i32p, C.c_int64,
                                      _u32p, _i32p, C.c_int64, C.c_int64, _i32p]
        L.cbsim_trans.argtypes = [C.c_int, C.c_int, C.c_int64, _u32p, _i32p, _u32p, _i32p]
        _lib = L
    return _lib


class SimDriver:
    def __init__(self, coop: bool = False):
        """coop: run the Frobenius (M = 1, homogeneous) and Horner bodies in their four-lanes-per-instance form
        (cb_coop.cuh) where the device would (2048 bits and up)"""
        self.coop = coop

    def fuchs(self, L, order, degree, valuation, n, batch, ode_m, ode_t, shared, res_m, res_t):
        return lib().cbsim_fuchs_batch(L, order, degree, valuation, n, batch, ode_m, ode_t, int(shared), res_m, res_t)

    def fuchs_range(self, L, order, degree, valuation, n_begin, n, batch, ode_m, ode_t, shared, res_m, res_t):
        return lib().cbsim_fuchs_batch_range(L, order, degree, valuation, n_begin, n, batch, ode_m, ode_t, int(shared),
                                             res_m, res_t)

    def fuchs_intop(self, L, order, degree, valuation, n, batch, ode_int, shared, res_m, res_t):
        return lib().cbsim_fuchs_intop_batch(L, order, degree, valuation, n, batch, ode_int, int(shared), res_m, res_t)

    def continuation(self, L, order, degree, n, batch, path_len, ode_m, ode_t, path_m, path_t, y_m, y_t):
        return lib().cbsim_continuation(L, order, degree, n, batch, path_len, ode_m, ode_t, path_m, path_t, y_m, y_t)

    def radius(self, L, order, degree, batch, ode_m, ode_t, shared, n, o_m, o_t, n_done):
        return lib().cbsim_radius(L, order, degree, batch, ode_m, ode_t, int(shared), n, o_m, o_t, n_done)

    def truncation(self, L, batch, e_m, e_t, a_m, a_t, bits, n_m, n_t, n_out):
        return lib().cbsim_truncation(L, batch, e_m, e_t, a_m, a_t, bits, n_m, n_t, n_out)

    def continuation_series(self, L, order, degree, n, batch, path_len, ode_m, ode_t, path_m, path_t, y_m, y_t, r_m, r_t):
        return lib().cbsim_continuation_series(L, order, degree, n, batch, path_len, ode_m, ode_t, path_m, path_t, y_m, y_t,
                                               r_m, r_t)

    def ode_shift(self, L, order, degree, batch, ode_m, ode_t, shared, a_m, a_t, shared_a, o_m, o_t):
        return lib().cbsim_ode_shift(L, order, degree, batch, ode_m, ode_t, int(shared), a_m, a_t, int(shared_a), o_m, o_t)

    def taylor_shift(self, L, length, batch, f_m, f_t, a_m, a_t, shared_a):
        return lib().cbsim_taylor_shift(L, length, batch, f_m, f_t, a_m, a_t, int(shared_a))

    def frobenius1(self, L, order, degree, valuation, n, batch, ode_m, ode_t, shared, rho_m, rho_t,
                   shared_rho, res_m, res_t, rhs_m=None, rhs_t=None, rhs_len=0, shared_rhs=False):
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        if self.coop and L >= 64 and rhs_len == 0:
            return lib().cbsim_frobenius1_coop(L, order, degree, valuation, n, batch, ode_m, ode_t, int(shared), rho_m, rho_t,
                                               int(shared_rho), res_m, res_t)
        return lib().cbsim_frobenius1_batch(L, order, degree, valuation, n, batch, ode_m, ode_t, int(shared),
                                            rho_m, rho_t, int(shared_rho), vp(rhs_m), vp(rhs_t), rhs_len,
                                            int(shared_rhs), res_m, res_t)

    def frobenius_general(self, L, order, degree, valuation, n, batch, ode_m, ode_t, shared, rho_m, rho_t,
                          shared_rho, mul, M, gens_m, gens_t):
        return lib().cbsim_frobenius_general_batch(L, order, degree, valuation, n, batch, ode_m, ode_t, int(shared),
                                                   rho_m, rho_t, int(shared_rho), mul, M, gens_m, gens_t)

    def solution_op(self, which, L, batch, mul, M, cap, nu, rho_m, rho_t, gens_m, gens_t, f_m, f_t, fcap):
        return lib().cbsim_solution_op(which, L, batch, mul, M, cap, nu, rho_m, rho_t, gens_m, gens_t, f_m, f_t, fcap)

    def indicial_polynomial(self, L, order, degree, valuation, batch, ode_m, ode_t, shared, nu, shift, out_m, out_t):
        return lib().cbsim_indicial_polynomial(L, order, degree, valuation, batch, ode_m, ode_t, int(shared), nu, shift,
                                               out_m, out_t)

    def solution_evaluate(self, L, M, cap, npts, g_m, g_t, rho_m, rho_t, pow_mode, pow_e, p_m, p_t, o_m, o_t):
        return lib().cbsim_solution_evaluate(L, M, cap, npts, g_m, g_t, rho_m, rho_t, pow_mode, pow_e, p_m, p_t, o_m, o_t)

    def trans(self, which, L, n, z_m, z_t, x_m, x_t):
        return lib().cbsim_trans(which, L, n, z_m, z_t, x_m, x_t)

    def ode_apply(self, L, order, degree, batch, ode_m, ode_t, shared, in_m, in_t, len_in, o_m, o_t, out_len, deg, solved):
        return lib().cbsim_ode_apply(L, order, degree, batch, ode_m, ode_t, int(shared), in_m, in_t, len_in, o_m, o_t,
                                     out_len, deg, solved)

    def horner(self, L, length, npts, nder, f_m, f_t, shared_f, p_m, p_t, o_m, o_t):
        if self.coop and L >= 64:
            return lib().cbsim_horner_coop(L, length, npts, nder, f_m, f_t, int(shared_f), p_m, p_t, o_m, o_t)
        return lib().cbsim_horner_batch(L, length, npts, nder, f_m, f_t, int(shared_f), p_m, p_t, o_m, o_t)

    def ew(self, op, L, n, z_m, z_t, x_m, x_t, y_m, y_t, k):
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        return lib().cbsim_ew(op, L, n, z_m, z_t, x_m, x_t, vp(y_m), vp(y_t), vp(k))


def path_count(i: int) -> int:
    """statistics of the simulated arithmetic (cb_arith.cuh CB_COUNT): e.g. 6 = two-row products of integer-like operands"""
    return int(lib().cbsim_path_count(i))
