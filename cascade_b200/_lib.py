"""Loader of the CUDA library ``libcascade_b200.so`` (C ABI: ``include/cascade_b200.h``).

There is deliberately no fallback: if the library is missing or no CUDA device is usable, every
compute call raises.  Build with ``python -c 'import __graft_entry__ as g; g.build()'`` or
``make -C cascade_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CB200_LIB") or os.path.join(_HERE, "libcascade_b200.so")  # CB200_LIB: a development build
_lib = None

EINVAL, EVALUATION, ERANGE, ENOMEM = -1, -2, -3, -4

_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_vp = C.c_void_p

SYMBOLS = [
    "cb200_version", "cb200_device_count", "cb200_kernel_launches", "cb200_debug_rr_stats",
    "cb200_fuchs_workspace_bytes", "cb200_frobenius_workspace_bytes",
    "cb200_horner_workspace_bytes", "cb200_ew_workspace_bytes",
    "cb200_fuchs_batch_dev", "cb200_fuchs_batch_host", "cb200_fuchs_batch_range_dev",
    "cb200_host_alloc", "cb200_host_free", "cb200_host_cache_release",
    "cb200_fuchs_intop_batch_dev", "cb200_fuchs_intop_batch_host",
    "cb200_continuation_workspace_bytes", "cb200_continuation_dev", "cb200_continuation_host",
    "cb200_continuation_series_workspace_bytes", "cb200_continuation_series_dev", "cb200_continuation_series_host",
    "cb200_radius_workspace_bytes", "cb200_radius_of_convergence_dev", "cb200_radius_of_convergence_host",
    "cb200_truncation_workspace_bytes", "cb200_truncation_order_dev", "cb200_truncation_order_host",
    "cb200_ode_shift_workspace_bytes", "cb200_ode_shift_dev", "cb200_ode_shift_host",
    "cb200_taylor_shift_workspace_bytes", "cb200_taylor_shift_dev", "cb200_taylor_shift_host",
    "cb200_frobenius1_batch_dev", "cb200_frobenius1_batch_host",
    "cb200_frobenius1_rhs_batch_dev", "cb200_frobenius1_rhs_batch_host",
    "cb200_horner_batch_dev", "cb200_horner_batch_host",
    "cb200_ew_dev", "cb200_ew_host",
    "cb200_frobenius_general_workspace_bytes", "cb200_frobenius_general_batch_dev",
    "cb200_frobenius_general_batch_host",
    "cb200_solution_op_workspace_bytes", "cb200_solution_op_dev", "cb200_solution_op_host",
    "cb200_indicial_polynomial_dev", "cb200_indicial_polynomial_host",
    "cb200_evaluate_workspace_bytes", "cb200_solution_evaluate_dev", "cb200_solution_evaluate_host",
    "cb200_trans_dev", "cb200_trans_host",
    "cb200_ode_apply_workspace_bytes", "cb200_ode_apply_dev", "cb200_ode_apply_host",
]


class CascadeB200Error(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CascadeB200Error(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run __graft_entry__.build()); cascade_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.cb200_version.restype = C.c_char_p
    L.cb200_device_count.restype = C.c_int
    L.cb200_kernel_launches.restype = C.c_uint64
    for name in ("cb200_fuchs_workspace_bytes",):
        getattr(L, name).restype = C.c_size_t
        getattr(L, name).argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int]
    for name in ("cb200_frobenius_workspace_bytes", "cb200_horner_workspace_bytes", "cb200_ew_workspace_bytes"):
        getattr(L, name).restype = C.c_size_t
        getattr(L, name).argtypes = [C.c_int, C.c_int64]
    L.cb200_fuchs_batch_dev.argtypes = [C.c_int] * 4 + [C.c_int64] * 2 + [_vp, _vp, C.c_int, _vp, _vp, _vp, C.c_size_t, _vp]
    L.cb200_fuchs_batch_range_dev.argtypes = [C.c_int] * 4 + [C.c_int64] * 3 + [_vp, _vp, C.c_int, _vp, _vp, _vp, C.c_size_t, _vp]
    L.cb200_host_alloc.restype = C.c_void_p
    L.cb200_host_alloc.argtypes = [C.c_size_t]
    L.cb200_host_free.argtypes = [C.c_void_p]
    L.cb200_host_cache_release.restype = None
    L.cb200_fuchs_batch_host.argtypes = [C.c_int] * 4 + [C.c_int64] * 2 + [_u32p, _i32p, C.c_int, _u32p, _i32p, C.c_int]
    _i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
    L.cb200_fuchs_intop_batch_dev.argtypes = [C.c_int] * 4 + [C.c_int64] * 2 + [_vp, C.c_int, _vp, _vp, _vp]
    L.cb200_fuchs_intop_batch_host.argtypes = [C.c_int] * 4 + [C.c_int64] * 2 + [_i64p, C.c_int, _u32p, _i32p, C.c_int]
    L.cb200_continuation_workspace_bytes.restype = C.c_size_t
    L.cb200_continuation_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64]
    L.cb200_continuation_dev.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp,
                                         _vp, _vp, _vp, _vp, C.c_size_t, _vp]
    L.cb200_continuation_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, _u32p, _i32p,
                                          _u32p, _i32p, _u32p, _i32p, C.c_int]
    L.cb200_continuation_series_workspace_bytes.restype = C.c_size_t
    L.cb200_continuation_series_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64]
    L.cb200_continuation_series_dev.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp,
                                                _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]
    L.cb200_continuation_series_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, _u32p, _i32p,
                                                 _u32p, _i32p, _u32p, _i32p, _u32p, _i32p, C.c_int]
    _i64p_ = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
    L.cb200_radius_workspace_bytes.restype = C.c_size_t
    L.cb200_radius_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int64]
    L.cb200_radius_of_convergence_dev.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _vp, _vp, C.c_int, C.c_int64, _vp, _vp, _vp,
                                                  _vp, C.c_size_t, _vp]
    L.cb200_radius_of_convergence_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _u32p, _i32p, C.c_int, C.c_int64, _u32p,
                                                   _i32p, _i32p, C.c_int]
    L.cb200_truncation_workspace_bytes.restype = C.c_size_t
    L.cb200_truncation_workspace_bytes.argtypes = [C.c_int, C.c_int64]
    L.cb200_truncation_order_dev.argtypes = [C.c_int, C.c_int64, _vp, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_size_t, _vp]
    L.cb200_truncation_order_host.argtypes = [C.c_int, C.c_int64, _u32p, _i32p, _u32p, _i32p, C.c_int64, _u32p, _i32p, _i64p_,
                                              C.c_int]
    L.cb200_ode_shift_workspace_bytes.restype = C.c_size_t
    L.cb200_ode_shift_workspace_bytes.argtypes = [C.c_int, C.c_int64]
    L.cb200_ode_shift_dev.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp,
                                      C.c_size_t, _vp]
    L.cb200_ode_shift_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _u32p, _i32p, C.c_int, _u32p, _i32p, C.c_int,
                                       _u32p, _i32p, C.c_int]
    L.cb200_taylor_shift_workspace_bytes.restype = C.c_size_t
    L.cb200_taylor_shift_workspace_bytes.argtypes = [C.c_int, C.c_int64, C.c_int64]
    L.cb200_taylor_shift_dev.argtypes = [C.c_int, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_size_t, _vp]
    L.cb200_taylor_shift_host.argtypes = [C.c_int, C.c_int64, C.c_int64, _u32p, _i32p, _u32p, _i32p, C.c_int, C.c_int]
    L.cb200_frobenius1_batch_dev.argtypes = [C.c_int] * 4 + [C.c_int64] * 2 + [_vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp, C.c_size_t, _vp]
    L.cb200_frobenius1_batch_host.argtypes = [C.c_int] * 4 + [C.c_int64] * 2 + [_u32p, _i32p, C.c_int, _u32p, _i32p, C.c_int, _u32p, _i32p, C.c_int]
    L.cb200_frobenius1_rhs_batch_dev.argtypes = [C.c_int] * 4 + [C.c_int64] * 2 + [_vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, C.c_int64, C.c_int, _vp, _vp, _vp, C.c_size_t, _vp]
    L.cb200_frobenius1_rhs_batch_host.argtypes = [C.c_int] * 4 + [C.c_int64] * 2 + [_u32p, _i32p, C.c_int, _u32p, _i32p, C.c_int, _vp, _vp, C.c_int64, C.c_int, _u32p, _i32p, C.c_int]
    L.cb200_horner_batch_dev.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]
    L.cb200_horner_batch_host.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int, _u32p, _i32p, C.c_int, _u32p, _i32p, _u32p, _i32p, C.c_int]
    L.cb200_ew_dev.argtypes = [C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]
    L.cb200_ew_host.argtypes = [C.c_int, C.c_int, C.c_int64, _u32p, _i32p, _u32p, _i32p, _vp, _vp, _vp, C.c_int]
    L.cb200_frobenius_general_workspace_bytes.restype = C.c_size_t
    L.cb200_frobenius_general_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64]
    L.cb200_frobenius_general_batch_dev.argtypes = ([C.c_int] * 4 + [C.c_int64] * 2 + [_vp, _vp, C.c_int, _vp, _vp, C.c_int,
                                                    C.c_int, C.c_int, _vp, _vp, _vp, C.c_size_t, _vp])
    L.cb200_frobenius_general_batch_host.argtypes = ([C.c_int] * 4 + [C.c_int64] * 2 + [_u32p, _i32p, C.c_int, _u32p, _i32p,
                                                     C.c_int, C.c_int, C.c_int, _u32p, _i32p, C.c_int])
    L.cb200_solution_op_workspace_bytes.restype = C.c_size_t
    L.cb200_solution_op_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int64]
    L.cb200_solution_op_dev.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64, _vp, _vp, _vp,
                                        _vp, _vp, _vp, C.c_int64, _vp, C.c_size_t, _vp]
    L.cb200_solution_op_host.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64, _u32p, _i32p,
                                         _u32p, _i32p, _u32p, _i32p, C.c_int64, C.c_int]
    L.cb200_indicial_polynomial_dev.argtypes = [C.c_int] * 4 + [C.c_int64, _vp, _vp, C.c_int, C.c_int64, C.c_int64, _vp, _vp,
                                                                _vp, C.c_size_t, _vp]
    L.cb200_indicial_polynomial_host.argtypes = [C.c_int] * 4 + [C.c_int64, _u32p, _i32p, C.c_int, C.c_int64, C.c_int64,
                                                                 _u32p, _i32p, C.c_int]
    L.cb200_evaluate_workspace_bytes.restype = C.c_size_t
    L.cb200_evaluate_workspace_bytes.argtypes = [C.c_int, C.c_int64]
    L.cb200_solution_evaluate_dev.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, C.c_uint64,
                                              _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]
    L.cb200_solution_evaluate_host.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, _u32p, _i32p, _u32p, _i32p, C.c_int,
                                               C.c_uint64, _u32p, _i32p, _u32p, _i32p, C.c_int]
    L.cb200_trans_dev.argtypes = [C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]
    L.cb200_trans_host.argtypes = [C.c_int, C.c_int, C.c_int64, _u32p, _i32p, _u32p, _i32p, C.c_int]
    L.cb200_ode_apply_workspace_bytes.restype = C.c_size_t
    L.cb200_ode_apply_workspace_bytes.argtypes = [C.c_int, C.c_int64, C.c_int64]
    L.cb200_ode_apply_dev.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _vp, _vp, C.c_int, _vp, _vp, C.c_int64, _vp, _vp,
                                      C.c_int64, C.c_int64, _vp, _vp, C.c_size_t, _vp]
    L.cb200_ode_apply_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _u32p, _i32p, C.c_int, _u32p, _i32p, C.c_int64,
                                       _u32p, _i32p, C.c_int64, C.c_int64, _i32p, C.c_int]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    names = {EINVAL: "invalid argument", EVALUATION: "valuation >= 0 (acb_ode_solve_fuchs is undefined there)",
             ERANGE: "integer factor overflows int64", ENOMEM: "workspace too small"}
    if rc < 0:
        raise CascadeB200Error(f"{what}: {names.get(rc, rc)}")
    raise CascadeB200Error(f"{what}: CUDA error {rc} (no usable device? cascade_b200 has no CPU fallback)")
