"""ctypes binding of libmfmg_b200.so (the C ABI declared in include/mfmg_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmfmg_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NOT_IMPLEMENTED, ERR_SINGULAR, ERR_NCCL, ERR_NOT_CONVERGED = range(7)


class MfmgError(RuntimeError):
    """std::runtime_error of the reference (ASSERT_THROW, include/mfmg/common/exceptions.hpp:59-63)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[mfmgb status {code}] {msg}")
        self.code = code


class NotImplementedExc(MfmgError, NotImplementedError):
    """mfmg::NotImplementedExc (include/mfmg/common/exceptions.hpp:65-84)."""


class NoConvergence(MfmgError):
    """dealii::SolverControl::NoConvergence."""

    def __init__(self, code, msg, iterations=None, history=None):
        super().__init__(code, msg)
        self.iterations = iterations
        self.history = history


_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_dbl = ctypes.c_double
_pp = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes)
SIGNATURES = {
    "mfmgb_ctx_create": (_int, [_int, _vp, _pp]),
    "mfmgb_ctx_destroy": (_int, [_vp]),
    "mfmgb_ctx_synchronize": (_int, [_vp]),
    "mfmgb_ctx_stream": (_vp, [_vp]),
    "mfmgb_last_error": (ctypes.c_char_p, [_vp]),
    "mfmgb_ctx_launch_count": (_i64, [_vp]),
    "mfmgb_version": (ctypes.c_char_p, []),
    "mfmgb_dev_malloc": (_int, [_vp, _i64, _pp]),
    "mfmgb_dev_free": (_int, [_vp, _vp]),
    "mfmgb_dev_upload": (_int, [_vp, _vp, _vp, _i64]),
    "mfmgb_dev_download": (_int, [_vp, _vp, _vp, _i64]),
    "mfmgb_vec_alloc": (_int, [_vp, _i64, _pp]),
    "mfmgb_vec_free": (_int, [_vp, _vp]),
    "mfmgb_vec_upload": (_int, [_vp, _vp, _vp, _i64]),
    "mfmgb_vec_download": (_int, [_vp, _vp, _vp, _i64]),
    "mfmgb_vec_fill": (_int, [_vp, _vp, _dbl, _i64]),
    "mfmgb_vec_copy": (_int, [_vp, _vp, _vp, _i64]),
    "mfmgb_vec_axpy": (_int, [_vp, _vp, _dbl, _vp, _i64]),
    "mfmgb_vec_dot": (_int, [_vp, _vp, _vp, _i64, ctypes.POINTER(_dbl)]),
    "mfmgb_csr_upload": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _pp]),
    "mfmgb_csr_upload_i32": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _pp]),
    "mfmgb_csr_adopt_device": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _pp]),
    "mfmgb_csr_destroy": (_int, [_vp, _vp]),
    "mfmgb_csr_info": (_int, [_vp, ctypes.POINTER(_i64), ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    "mfmgb_csr_device_arrays": (_int, [_vp, _pp, _pp, _pp, ctypes.POINTER(_int)]),
    "mfmgb_csr_download": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "mfmgb_csr_transpose": (_int, [_vp, _vp, _pp]),
    "mfmgb_csr_set_lanes_per_row": (_int, [_vp, _int]),
    "mfmgb_csr_get_lanes_per_row": (_int, [_vp]),
    "mfmgb_csr_set_kernel": (_int, [_vp, _int]),
    "mfmgb_csr_get_kernel": (_int, [_vp]),
    "mfmgb_spmv": (_int, [_vp, _vp, _vp, _vp]),
    "mfmgb_residual_neg": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "mfmgb_restrict": (_int, [_vp, _vp, _vp, _vp]),
    "mfmgb_prolong_correct": (_int, [_vp, _vp, _vp, _vp]),
    "mfmgb_jacobi_setup": (_int, [_vp, _vp, _dbl, _pp]),
    "mfmgb_jacobi_setup_diag": (_int, [_vp, _vp, _i64, _dbl, _pp]),
    "mfmgb_jacobi_destroy": (_int, [_vp, _vp]),
    "mfmgb_jacobi_inv_diag": (_vp, [_vp]),
    "mfmgb_jacobi_apply": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "mfmgb_jacobi_apply_oop": (_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mfmgb_jacobi_apply_residual": (_int, [_vp, _vp, _vp, _vp]),
    "mfmgb_jacobi_apply_zero_guess": (_int, [_vp, _vp, _vp, _vp]),
    "mfmgb_dense_factor": (_int, [_vp, _vp, _pp]),
    "mfmgb_dense_destroy": (_int, [_vp, _vp]),
    "mfmgb_dense_solve": (_int, [_vp, _vp, _vp, _vp]),
    "mfmgb_dense_size": (_i64, [_vp]),
    "mfmgb_dense_num_swaps": (_i64, [_vp]),
    "mfmgb_dense_solve_mode": (_int, [_vp, ctypes.POINTER(_dbl)]),
    "mfmgb_mf_laplace_create": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp, _pp]),
    "mfmgb_mf_laplace_create_slab": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp, _i64, _i64, _pp]),
    "mfmgb_mf_vector_size": (_i64, [_vp]),
    "mfmgb_mf_kernel": (_int, [_vp]),
    "mfmgb_mf_destroy": (_int, [_vp, _vp]),
    "mfmgb_mf_size": (_i64, [_vp]),
    "mfmgb_mf_apply": (_int, [_vp, _vp, _vp, _vp]),
    "mfmgb_mf_diagonal": (_int, [_vp, _vp, _vp]),
    "mfmgb_hierarchy_create": (_int, [_vp, _int, _int, _int, _dbl, _pp]),
    "mfmgb_hierarchy_set_operator": (_int, [_vp, _int, _vp]),
    "mfmgb_hierarchy_set_mf_operator": (_int, [_vp, _vp]),
    "mfmgb_hierarchy_set_restrictor": (_int, [_vp, _int, _vp, _vp]),
    "mfmgb_hierarchy_set_smoother_chebyshev": (_int, [_vp, _int, _dbl, _dbl, _int]),
    "mfmgb_hierarchy_chebyshev_info": (_int, [_vp, _int, _vp]),
    "mfmgb_hierarchy_finalize": (_int, [_vp, _vp]),
    "mfmgb_hierarchy_destroy": (_int, [_vp, _vp]),
    "mfmgb_vcycle": (_int, [_vp, _vp, _vp, _vp]),
    "mfmgb_hierarchy_apply": (_int, [_vp, _vp, _vp, _vp, _int]),
    "mfmgb_vcycle_host": (_int, [_vp, _vp, _vp, _vp]),
    "mfmgb_vcycle_host_batch": (_int, [_vp, _vp, _int, _vp, _vp]),
    "mfmgb_vcycle_profile": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "mfmgb_vcycle_timeline": (_int, [_vp, _vp, _vp, _vp, _int, ctypes.c_char_p, _int, _vp, _int, ctypes.POINTER(_int)]),
    "mfmgb_hierarchy_use_graph": (_int, [_vp, _int]),
    "mfmgb_hierarchy_launches_per_cycle": (_int, [_vp]),
    "mfmgb_pcg": (_int, [_vp, _vp, _vp, _vp, _vp, _dbl, _int, ctypes.POINTER(_int), _vp]),
    "mfmgb_pcg_host": (_int, [_vp, _vp, _vp, _vp, _vp, _dbl, _int, ctypes.POINTER(_int), _vp]),
    "mfmgb_comm_unique_id": (_int, [ctypes.c_char_p]),
    "mfmgb_comm_init": (_int, [_vp, ctypes.c_char_p, _int, _int]),
    "mfmgb_comm_finalize": (_int, [_vp]),
    "mfmgb_comm_transport": (ctypes.c_char_p, [_vp]),
    "mfmgb_comm_check": (_int, [_vp]),
    "mfmgb_comm_rank": (_int, [_vp]),
    "mfmgb_comm_size": (_int, [_vp]),
    "mfmgb_halo_create": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _pp]),
    "mfmgb_halo_destroy": (_int, [_vp, _vp]),
    "mfmgb_halo_exchange": (_int, [_vp, _vp, _vp]),
    "mfmgb_allreduce_sum": (_int, [_vp, _vp, _int]),
    "mfmgb_hierarchy_set_halo": (_int, [_vp, _int, _vp, _i64, _i64]),
    "mfmgb_hierarchy_set_coarse_offsets": (_int, [_vp, _vp, _int]),
    "mfmgb_hierarchy_vector_size": (_i64, [_vp, _int]),
    "mfmgb_tunable_set": (_int, [ctypes.c_char_p, ctypes.c_longlong]),
    "mfmgb_tunable_get": (ctypes.c_longlong, [ctypes.c_char_p]),
    "mfmgb_coarse_dd_create": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _pp]),
    "mfmgb_coarse_dd_destroy": (_int, [_vp, _vp]),
    "mfmgb_coarse_dd_solve": (_int, [_vp, _vp, _vp, _vp]),
    "mfmgb_hierarchy_set_coarse_dd": (_int, [_vp, _vp]),
    "mfmgb_hierarchy_set_restrict_split": (_int, [_vp, _int, _i64]),
    "mfmgb_hierarchy_set_restrict_no_halo": (_int, [_vp, _vp, _int, _vp]),
}

_LIB = None


def _preload_bundled_nccl() -> None:
    """libmfmg_b200.so needs libnccl.so.2.  When this module is loaded before torch, the dynamic linker would pick
    the system NCCL, and torch (which needs the newer NCCL bundled with its wheel) would then fail to import in the
    same process.  Loading the bundled copy first makes both share it; without one, the system NCCL serves."""
    import importlib.util

    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for base in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
        cand = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            ctypes.CDLL(cand, mode=ctypes.RTLD_GLOBAL)
            return


def load() -> ctypes.CDLL:
    """Load the CUDA library.  Raises if it has not been built (python -m mfmg_b200.build)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the sm_100a extension has not been built "
                "(run `python -m mfmg_b200.build` or __graft_entry__.build()); there is no CPU fallback")
        _preload_bundled_nccl()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def check(ctx, rc: int) -> None:
    if rc == OK:
        return
    msg = load().mfmgb_last_error(ctx)
    msg = msg.decode() if msg else "unknown error"
    if rc == ERR_NOT_IMPLEMENTED:
        raise NotImplementedExc(rc, msg)
    if rc == ERR_NOT_CONVERGED:
        raise NoConvergence(rc, msg)
    raise MfmgError(rc, msg)
