"""ctypes binding of libb200arnoldi.so - every symbol include/b200arnoldi.h declares.

The product has NO CPU fallback: if the shared library is missing, this module
raises at import of the first symbol; compute entry points fail with
``B200Error`` when no sm_100 device is present.
"""

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200arnoldi.so")

# enums of include/b200arnoldi.h
F64, C64 = 0, 1
WHICH = {"LM": 0, "LR": 1, "SR": 2, "LI": 3, "SI": 4}
INIT_NONE, INIT_RAND, INIT_KEEP = 0, 1, 2
KERNEL_KINDS = ("spmv", "cgs_dots", "cgs_update", "cgs_finish", "rotate", "fill", "cgs_sweep", "xchg")
OK, ERR_ARGUMENT, ERR_DIMENSION, ERR_CUDA, ERR_NCCL, ERR_OOM, ERR_QR, ERR_INTERNAL, ERR_CALLBACK, ERR_SOLVE = (
    0, -1, -2, -3, -4, -5, -6, -7, -8, -9,
)
SOLVE_CG = 0


class Stats(C.Structure):
    _fields_ = [
        ("matvecs", C.c_int64),
        ("passes", C.c_int64),
        ("second_passes", C.c_int64),
        ("breakdowns", C.c_int64),
        ("launches", C.c_int64),
        ("bytes", C.c_double),
    ]


class Params(C.Structure):
    _fields_ = [
        ("nev", C.c_int32),
        ("which", C.c_int32),
        ("tol", C.c_double),
        ("mindim", C.c_int32),
        ("maxdim", C.c_int32),
        ("restarts", C.c_int32),
        ("start_from", C.c_int32),
        ("initialize", C.c_int32),
        ("seed", C.c_uint64),
    ]


class HistoryC(C.Structure):
    _fields_ = [
        ("mvproducts", C.c_int64),
        ("nconverged", C.c_int32),
        ("converged", C.c_int32),
        ("nev", C.c_int32),
        ("restarts", C.c_int32),
        ("stats", Stats),
        ("ms_expand", C.c_double),
        ("ms_rotate", C.c_double),
        ("ms_small", C.c_double),
    ]


MATVEC_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)

_vp, _i, _i64, _u64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
_pvp, _pi, _pi64, _pd = C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_double)

# name -> (restype, argtypes); mirrors include/b200arnoldi.h one to one
SIGNATURES = {
    "b2a_version": (_i, []),
    "b2a_last_error": (C.c_char_p, []),
    "b2a_ctx_create": (_i, [_i, _pvp]),
    "b2a_ctx_create_dist": (_i, [_i, _i, _i, _vp, _pvp]),
    "b2a_nccl_unique_id": (_i, [_vp]),
    "b2a_ctx_destroy": (_i, [_vp]),
    "b2a_ctx_stream": (_i, [_vp, _pvp]),
    "b2a_ctx_sync": (_i, [_vp]),
    "b2a_ctx_rank": (_i, [_vp, _pi, _pi]),
    "b2a_ctx_launch_count": (_i, [_vp, _pi64]),
    "b2a_ctx_profile_enable": (_i, [_vp, _i]),
    "b2a_ctx_profile_get": (_i, [_vp, _i, _pi64, _pd, _pd]),
    "b2a_csr_create": (_i, [_vp, _i, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _i, _i, _pvp]),
    "b2a_csr_create_device": (_i, [_vp, _i, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _pvp]),
    "b2a_csc_create": (_i, [_vp, _i, _i64, _i64, _vp, _vp, _vp, _i, _i, _i, _pvp]),
    "b2a_op_from_callback": (_i, [_vp, _i, _i64, _i64, MATVEC_FN, _vp, _pvp]),
    "b2a_op_shift_invert": (_i, [_vp, _vp, _d, _d, _i, _d, _i, _pvp]),
    "b2a_op_solve_stats": (_i, [_vp, _pi64, _pi64, _pd]),
    "b2a_op_destroy": (_i, [_vp]),
    "b2a_op_bytes": (_i, [_vp, _pd]),
    "b2a_ws_create": (_i, [_vp, _i, _i64, _i64, _i64, _i, _pvp]),
    "b2a_ws_destroy": (_i, [_vp]),
    "b2a_ws_set_col": (_i, [_vp, _i, _vp]),
    "b2a_ws_set_col_device": (_i, [_vp, _i, _vp]),
    "b2a_ws_get_cols": (_i, [_vp, _i, _i, _vp, _i64]),
    "b2a_ws_comm_mode": (_i, [_vp, _pi]),
    "b2a_ws_debug_sweep_trace": (_i, [_vp, _vp, _i, _pi]),
    "b2a_ws_col_ptr": (_i, [_vp, _i, _pvp, _pi64]),
    "b2a_ws_host_arrays": (_i, [_vp, _pvp, _pi, _pvp, _pi]),
    "b2a_reinitialize": (_i, [_vp, _i, _i, _u64, _pi]),
    "b2a_orthogonalize": (_i, [_vp, _i, _vp, _pi]),
    "b2a_iterate_arnoldi": (_i, [_vp, _vp, _i, _i, _u64, _vp, _i, C.POINTER(Stats)]),
    "b2a_rotate_basis": (_i, [_vp, _i, _i, _i, _vp, _i, C.POINTER(Stats)]),
    "b2a_rotate_final": (_i, [_vp, _i, _vp, _i, C.POINTER(Stats)]),
    "b2a_basis_times": (_i, [_vp, _i, _vp, _i, _vp, _i64]),
    "b2a_ws_matvec": (_i, [_vp, _vp, _i, _i]),
    "b2a_ws_nrm2": (_i, [_vp, _i, _pd]),
    "b2a_ws_gemv_c": (_i, [_vp, _i, _i, _vp]),
    "b2a_ws_gemv_n_sub": (_i, [_vp, _i, _i, _vp]),
    "b2a_ws_scal_div": (_i, [_vp, _i, _d]),
    "b2a_ws_copy_col": (_i, [_vp, _i, _i]),
    "b2a_partialschur": (_i, [_vp, _vp, C.POINTER(Params), C.POINTER(HistoryC), _vp]),
    "b2a_host_local_schurfact": (_i, [_i, _vp, _i, _i, _i, _i, _i, _vp, _i, _i]),
    "b2a_host_restart": (_i, [_i, _vp, _i, _vp, _i, _i, _i, _i, _d, _i, _i, _pi, _pi, _pi, _vp, _vp]),
    "b2a_host_sortschur": (_i, [_i, _vp, _i, _vp, _i, _i, _i, _i]),
    "b2a_host_col_block_plan": (_i, [_i, _i64, _d, _d, _pi]),
    "b2a_host_owner_group_plan": (_i, [_i, _i64, _i, _i, _pi, _pi, _pi]),
    "b2a_host_givens": (_i, [_i, _vp, _vp, _vp, _vp, _vp]),
}


class B200Error(RuntimeError):
    """Raised for CUDA / NCCL / internal failures (status <= -3)."""

    def __init__(self, status, message):
        super().__init__(f"libb200arnoldi status {status}: {message}")
        self.status = status


class DimensionMismatch(ValueError):
    """Mirror of Julia's DimensionMismatch (src/run.jl:110)."""


_lib = None


def lib():
    """Load the shared library (once).  No fallback: a missing library is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing - build it with `python arnoldimethod.jl_b200/build.py` "
                "(there is no CPU fallback)"
            )
        try:
            import torch  # noqa: F401  (maps libnccl.so.2 so that dlopen() inside the library finds it)
        except Exception:
            pass
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    """Translate a status code into the reference's exception classes."""
    if status == OK:
        return
    msg = lib().b2a_last_error().decode("utf-8", "replace")
    if status == ERR_ARGUMENT:
        raise ValueError(msg)  # Julia ArgumentError
    if status == ERR_DIMENSION:
        raise DimensionMismatch(msg)
    raise B200Error(status, msg)
