"""ctypes binding of libb200ls.so -- exactly the entry points declared in include/b200ls.h.

This is the binding a PetIBM maintainer's C++ shim uses as well; Python is only
the harness language.  If the shared library is missing the import fails loudly: there is no
eager/CPU fallback in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200ls.so")

OK = 0
ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_NCCL, ERR_DIVERGED, ERR_MISMATCH, ERR_PARSE = -1, -2, -3, -4, -5, -6, -7
KSP_CG, KSP_BCGS, KSP_PREONLY = 0, 1, 2
PC_NONE, PC_JACOBI, PC_MG, PC_LU = 0, 1, 2, 3
NORM_NONE, NORM_PRECONDITIONED, NORM_UNPRECONDITIONED, NORM_NATURAL = 0, 1, 2, 3
REDUCE_P2P, REDUCE_NCCL = 0, 1
HALO_STORE, HALO_MEMCPY = 0, 1


class Options(C.Structure):
    _fields_ = [
        ("ksp_type", C.c_int),
        ("pc_type", C.c_int),
        ("norm_type", C.c_int),
        ("max_it", C.c_int),
        ("rtol", C.c_double),
        ("atol", C.c_double),
        ("divtol", C.c_double),
        ("check_every", C.c_int),
        ("variant", C.c_int),
        ("mg_levels", C.c_int),
        ("mg_smooth_its", C.c_int),
        ("mg_coarse_its", C.c_int),
    ]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); kept in one table so that tests can check it against the header
SIGNATURES = {
    "b200ls_version": (C.c_int, []),
    "b200ls_error_string": (C.c_char_p, [C.c_int]),
    "b200ls_device_count": (C.c_int, [_ip]),
    "b200ls_axis_from_subdomains": (C.c_int, [C.c_double, C.c_int, _dp, _ip, _dp, _dp, C.c_int, _ip]),
    "b200ls_default_options": (None, [C.POINTER(Options)]),
    "b200ls_parse_options": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(Options), C.c_char_p, C.c_size_t]),
    "b200ls_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "b200ls_destroy": (C.c_int, [_vp]),
    "b200ls_last_error": (C.c_char_p, [_vp]),
    "b200ls_set_options": (C.c_int, [_vp, C.POINTER(Options)]),
    "b200ls_get_options": (C.c_int, [_vp, C.POINTER(Options)]),
    "b200ls_set_tuning": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "b200ls_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "b200ls_comm_export": (C.c_int, [_vp, _vp]),
    "b200ls_comm_connect": (C.c_int, [_vp, _vp, C.c_int]),
    "b200ls_comm_disconnect": (C.c_int, [_vp]),
    "b200ls_nccl_unique_id": (C.c_int, [_vp]),
    "b200ls_nccl_init": (C.c_int, [_vp, _vp]),
    "b200ls_dmda_split": (C.c_int, [C.c_int64, C.c_int, _i64p]),
    "b200ls_repart_create": (C.c_int, [C.POINTER(_vp), C.c_int, _i64p, _ip, C.c_int]),
    "b200ls_repart_destroy": (C.c_int, [_vp]),
    "b200ls_repart_info": (C.c_int, [_vp, _i64p, _i64p, _i64p, _i64p, _i64p, _i64p, _ip]),
    "b200ls_repart_counts": (C.c_int, [_vp, _i64p, _i64p, _i64p, _i64p]),
    "b200ls_repart_unpack_slab": (C.c_int, [_vp, _dp, _dp]),
    "b200ls_repart_pack_slab": (C.c_int, [_vp, _dp, _dp]),
    "b200ls_repart_petsc_to_natural": (C.c_int, [_vp, C.c_int64, _i32p, _i32p]),
    "b200ls_repart_box_rows": (C.c_int, [_vp, _i64p]),
    "b200ls_repart_candidates": (C.c_int, [C.c_int, _i64p, C.c_int, _i64p, _ip, C.c_int, _ip]),
    "b200ls_set_poisson_stencil": (C.c_int, [_vp, C.c_int, _i64p, _ip, _dp, _dp, _dp, C.c_double, C.c_int64, C.c_int64]),
    "b200ls_verify_csr": (C.c_int, [_vp, C.c_int64, _i64p, _i32p, _dp, _dp]),
    "b200ls_verify_csr_rows": (C.c_int, [_vp, C.c_int64, _i64p, _i64p, _i32p, _dp, C.c_int, _dp]),
    "b200ls_set_csr": (C.c_int, [_vp, C.c_int64, _i64p, _i32p, _dp]),
    "b200ls_set_staggered": (C.c_int, [_vp, C.c_int, _i64p, _ip, C.c_int64, _i64p, _i32p, _dp]),
    "b200ls_staggered_coef_size": (C.c_int64, [C.c_int, _i64p]),
    "b200ls_staggered_analyze": (C.c_int, [C.c_int, _i64p, _ip, C.c_int64, _i64p, _i32p, _dp, _dp, _dp, _i64p, _i32p, _dp,
                                          C.c_char_p, C.c_size_t]),
    "b200ls_set_poisson_hybrid": (C.c_int, [_vp, C.c_int, _i64p, _ip, _dp, _dp, _dp, C.c_double, C.c_int64, _i64p, _i32p, _dp]),
    "b200ls_hybrid_analyze": (C.c_int, [C.c_int, _i64p, _ip, _dp, _dp, _dp, C.c_double, C.c_int64, _i64p, _i32p, _dp, _dp, _dp,
                                       _i64p, _i32p, _dp, C.c_char_p, C.c_size_t]),
    "b200ls_set_nullspace": (C.c_int, [_vp, C.c_int, C.c_int, _dp]),
    "b200ls_apply": (C.c_int, [_vp, _vp, _vp]),
    "b200ls_solve": (C.c_int, [_vp, _vp, _vp]),
    "b200ls_solve_device": (C.c_int, [_vp, _vp, _vp]),
    "b200ls_get_iters": (C.c_int, [_vp, _ip]),
    "b200ls_get_residual": (C.c_int, [_vp, _dp]),
    "b200ls_get_reason": (C.c_int, [_vp, _ip]),
    "b200ls_get_history": (C.c_int, [_vp, _dp, C.c_int, _ip]),
    "b200ls_velocity_size": (C.c_int, [_vp, _i64p, _i64p]),
    "b200ls_divergence_device": (C.c_int, [_vp, _vp, _vp]),
    "b200ls_gradient_device": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "b200ls_project_device": (C.c_int, [_vp, _vp, _vp, _vp]),
    "b200ls_ghosted_sizes": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "b200ls_convection_device": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "b200ls_ghosted_from_packed_device": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "b200ls_convection": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "b200ls_divergence": (C.c_int, [_vp, _vp, _vp]),
    "b200ls_gradient": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "b200ls_project": (C.c_int, [_vp, _vp, _vp, _vp]),
    "b200ls_get_timing": (C.c_int, [_vp, _dp, _dp, _i64p]),
    "b200ls_get_e2e_ms": (C.c_int, [_vp, _dp]),
    "b200ls_set_profile": (C.c_int, [_vp, C.c_int]),
    "b200ls_get_profile": (C.c_int, [_vp, C.c_int, _dp, _i64p]),
    "b200ls_time_kernel": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp]),
    "b200ls_set_trace": (C.c_int, [_vp, C.c_int]),
    "b200ls_get_trace": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.c_int, _ip]),
    "b200ls_stream": (_vp, [_vp]),
}

_lib = None


class B200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200ls error {code}: {msg}")
        self.code = code


def lib():
    """Load libb200ls.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m petibm_b200.build` "
            "(the B200 backend has no CPU or PyTorch fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the header and the library drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(code: int, handle=None):
    if code == OK:
        return
    L = lib()
    msg = L.b200ls_error_string(code).decode()
    if handle:
        detail = L.b200ls_last_error(handle).decode()
        if detail:
            msg += ": " + detail
    raise B200Error(code, msg)
