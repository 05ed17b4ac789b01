"""ctypes binding of libscmerge_b200.so (the C ABI declared in include/scmerge_b200.h).

There is deliberately no fallback: if the shared library is missing or no B200 is visible the
import / context creation raises.  Build the library with `python -c "import __graft_entry__ as g; g.build()"`
or `make -C scmerge_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscmerge_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NONFINITE, ERR_SINGULAR, ERR_UNSUPPORTED, ERR_NOCONV, ERR_COMM = 0, -1, -2, -3, -4, -5, -6, -7
MODE_EXACT, MODE_RANDOMIZED = 0, 1
MAX_K, MAX_BLOCK = 64, 112
MAX_TOPK_ROWS, MAX_FULL_EIG_N = 88, 4096  # include/scmerge_b200.h

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p)

# name -> argtypes; every function returns int unless listed in _RESTYPES
_vp, _i64, _i32, _int, _dbl, _u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_int, C.c_double, C.c_uint64
_pp = C.POINTER(C.c_void_p)
SIGNATURES = {
    "scm_last_error": [],
    "scm_version": [],
    "scm_device_pci_bus_id": [_int, C.c_char_p, _int],
    "scm_ctx_create": [_int, _vp, _pp],
    "scm_ctx_destroy": [_vp],
    "scm_ctx_set_allreduce": [_vp, ALLREDUCE_FN, _vp, _int],
    "scm_nccl_unique_id": [_vp],
    "scm_ctx_comm_init_nccl": [_vp, _vp, _int, _int],
    "scm_ctx_comm_info": [_vp, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int)],
    "scm_allreduce_sum": [_vp, _vp, _i64, _int],
    "scm_ctx_fit_info": [_vp, C.POINTER(_i32), C.POINTER(_dbl), C.POINTER(_i32), C.POINTER(_i32)],
    "scm_multi_create": [C.POINTER(_int), _int, _pp],
    "scm_multi_destroy": [_vp],
    "scm_multi_size": [_vp],
    "scm_multi_ctx": [_vp, _int],
    "scm_multi_ruv3_host": [_vp, _vp, _i64, _i64, _i64, _vp, _i32, _vp, _i32, _vp, _i32, _i32, _int, _dbl, _i32, _u64, _vp, _i32, _int,
                            _vp, _vp, _i64, _vp, _vp, _vp, _vp, C.POINTER(_i32), C.POINTER(_dbl)],
    "scm_ctx_set_conv_ranks": [_vp, _vp, _i32],
    "scm_ctx_sync": [_vp],
    "scm_ctx_launch_count": [_vp, C.POINTER(_i64)],
    "scm_ctx_release_scratch": [_vp],
    "scm_dev_alloc": [_vp, _i64, _pp],
    "scm_dev_free": [_vp, _vp],
    "scm_host_alloc_pinned": [_i64, _pp],
    "scm_host_free_pinned": [_vp],
    "scm_h2d": [_vp, _vp, _vp, _i64],
    "scm_d2h": [_vp, _vp, _vp, _i64],
    "scm_copy2d": [_vp, _vp, _i64, _vp, _i64, _i64, _i64, _int],
    "scm_transpose": [_vp, _vp, _i64, _i64, _i64, _vp, _i64],
    "scm_cosine_rows": [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _dbl],
    "scm_csc_to_dense": [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _dbl, _vp, _i64, _vp, _vp],
    "scm_csc_check": [_vp, _vp, _vp, _i64, _i64, _i64, C.POINTER(_i32)],
    "scm_csc_rownorm": [_vp, _vp, _vp, _i64, _int, _dbl, _vp, _vp],
    "scm_csc_seg_colmean": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _i32, _vp, _i64, _vp],
    "scm_csc_adjust": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _i64],
    "scm_csc_adjust_to_host": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _i64],
    "scm_adjust_to_host": [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _i64],
    "scm_seg_colsum": [_vp, _vp, _i64, _i64, _i64, _vp, _i32, _vp, _vp, _vp, _vp],
    "scm_seg_colmean": [_vp, _vp, _i64, _i64, _i64, _vp, _i32, _vp, _vp, _vp],
    "scm_group_colmean": [_vp, _vp, _vp, _i32, _i64, _i64, _vp],
    "scm_gram": [_vp, _vp, _i64, _i64, _i64, _vp, _vp],
    "scm_gram_residual": [_vp, _vp, _i64, _vp, _vp, _i32, _vp, _vp],
    "scm_eig_topk": [_vp, _vp, _i64, _i32, _i32, _int, _dbl, _i32, _u64, _vp, _vp, C.POINTER(_i32), C.POINTER(_dbl)],
    "scm_fullalpha": [_vp, _vp, _vp, _i32, _i64, _vp],
    "scm_coef_create": [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _pp],
    "scm_coef_destroy": [_vp, _vp],
    "scm_coef_from_factors": [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _pp],
    "scm_ruvg_fit": [_vp, _vp, _i64, _i64, _i64, _vp, _i32, _int, _dbl, _i32, _u64, _vp, _vp, _vp, _pp, C.POINTER(_i32),
                     C.POINTER(_dbl)],
    "scm_adjust": [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _i32, _vp, _i64, _vp],
    "scm_ruv3_fit": [_vp, _vp, _i64, _i64, _i64, _vp, _i32, _vp, _i32, _i32, _i32, _int, _dbl, _i32, _u64,
                     _vp, _vp, _vp, _vp, C.POINTER(_i32), C.POINTER(_dbl)],
    "scm_ruv3_fit_gram": [_vp, _vp, _i64, _i64, _i64, _vp, _i32, _vp, _i32, _i32, _i32, _int, _dbl, _i32, _u64,
                          _vp, _vp, _vp, _vp, C.POINTER(_i32), C.POINTER(_dbl), _vp, _vp],
    "scm_pca_scores": [_vp, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _i64, _i64, _vp, _vp, _dbl, _i32, C.POINTER(_i32), C.POINTER(_dbl)],
    "scm_silhouette": [_vp, _vp, _i64, _i32, _i64, _vp, _i32, _vp, _i32, _i64, _i64, C.POINTER(_dbl), C.POINTER(_dbl)],
    "scm_ruv3_host": [_vp, _vp, _i64, _i64, _i64, _vp, _i32, _vp, _i32, _vp, _i32, _i32, _int, _dbl, _i32, _u64, _vp, _i32, _int,
                      _vp, _vp, _i64, _vp, _vp, _vp, _vp, C.POINTER(_i32), C.POINTER(_dbl)],
    "scm_gather": [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64],
    "scm_group_median": [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _i32, _vp],
    "scm_sqdist_to_center": [_vp, _vp, _i64, _i64, _i64, _vp, _i32, _vp, _vp],
    "scm_pair_dist": [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _int, _vp, _i64],
    "scm_block_median": [_vp, _vp, _i64, _vp, _i32, _vp, _i32, _vp],
    "scm_centred_gram": [_vp, _vp, _i64, _i64, _i64, _vp, _vp, C.POINTER(_i64)],
    "scm_ls_coef": [_vp, _vp, _i64, _i64, _i64, _vp, _i32, _vp, _vp],
    "scm_rankk_subtract": [_vp, _vp, _i64, _i64, _i64, _vp, _i32, _vp, _vp, _i64],
    "scm_kmeans": [_vp, _vp, _i64, _i32, _i64, _i32, _i32, _i32, _u64, _vp, _vp, _vp, _vp, C.POINTER(_dbl), _vp],
    "scm_timers": [_vp, C.POINTER(_dbl), _int],
}
_RESTYPES = {"scm_last_error": C.c_char_p, "scm_multi_ctx": C.c_void_p}


class ScmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[scmerge_b200 {code}] {msg}")
        self.code = code


class NonFiniteError(ScmError, ValueError):
    """Y contains NA / Inf (the reference's stop(), R/fastRUVIII.R:52-60)."""


class SingularError(ScmError, ArithmeticError):
    """ac %*% t(ac) is singular (R's solve() error, R/fastRUVIII.R:99)."""


class NotConvergedWarning(UserWarning):
    pass


_lib = None


def lib():
    """Load the shared library once; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA library has not been built and scmerge_b200 has no CPU fallback. "
                "Run `make -C scmerge_b200/csrc` (needs nvcc, builds for sm_100a).")
        handle = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = handle
    return _lib


def check(rc: int) -> int:
    if rc == OK:
        return rc
    msg = lib().scm_last_error().decode("utf-8", "replace")
    if rc == ERR_NONFINITE:
        raise NonFiniteError(rc, msg)
    if rc == ERR_SINGULAR:
        raise SingularError(rc, msg)
    raise ScmError(rc, msg)
