"""ctypes binding of the C ABI declared in include/mixdir_b200.h.

The CUDA library is the product; there is no CPU fallback.  Importing this
module without the built library raises ImportError (build it with
`python -c "import __graft_entry__ as g; g.build()"` or `make -C mixdir_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmixdir_b200.so")

OK, UNDERFLOW, CUDA_ERROR, BAD_ARGUMENT = 0, 1, 2, 3
WARN_ELBO_DECREASED, WARN_DP_LAST_COMPONENT = 1, 2

# every symbol include/mixdir_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "mixdir_b200_ragged_size", "mixdir_b200_create", "mixdir_b200_create_from_r_matrix",
    "mixdir_b200_nccl_unique_id", "mixdir_b200_comm_init", "mixdir_b200_comm_destroy",
    "mixdir_b200_create_dist", "mixdir_b200_create_dist_from_r_matrix", "mixdir_b200_destroy", "mixdir_b200_release_cache",
    "mixdir_b200_n_rows", "mixdir_b200_max_code", "mixdir_b200_n_missing", "mixdir_b200_fit",
    "mixdir_b200_fit_batch",
    "mixdir_b200_predict", "mixdir_b200_feature_divergence", "mixdir_b200_estep",
    "mixdir_b200_last_timing",
    "mixdir_b200_set_engine", "mixdir_b200_set_pairing", "mixdir_b200_unit_plan", "mixdir_b200_unit_layout",
    "mixdir_b200_last_error",
    "mixdir_b200_abi_version",
]


class FitParams(C.Structure):
    _fields_ = [("n_latent", C.c_int), ("dp", C.c_int), ("alpha1", C.c_double),
                ("alpha2", C.c_double), ("beta", C.c_double), ("max_iter", C.c_int),
                ("epsilon", C.c_double), ("n_cat_bound", C.c_int)]


class Timing(C.Structure):
    _fields_ = [("fit_loop_ms", C.c_double), ("em_kernel_ms", C.c_double),
                ("iterations", C.c_int), ("em_launches", C.c_int), ("total_launches", C.c_int),
                ("engine", C.c_int), ("classes_per_cta", C.c_int), ("cluster_size", C.c_int),
                ("grid_ctas", C.c_int), ("smem_bytes", C.c_int), ("unit_elements", C.c_int),
                ("unit_pairs", C.c_int), ("unit_table_rows", C.c_int),
                ("final_pass_engine", C.c_int), ("final_pass_kernel_ms", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA engine is not built and mixdir_b200 has no CPU "
            "fallback.  Run `make -C mixdir_b200/csrc` (needs nvcc).")
    L = C.CDLL(LIB_PATH)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.mixdir_b200_ragged_size.restype = C.c_size_t
    L.mixdir_b200_ragged_size.argtypes = [C.c_int, ip, C.c_int]
    L.mixdir_b200_create.argtypes = [vp, C.c_int, C.c_int64, C.c_int, ip, C.c_int, ip, C.POINTER(vp)]
    L.mixdir_b200_create_from_r_matrix.argtypes = [vp, C.c_int64, C.c_int, ip, C.c_int, ip, C.POINTER(vp)]
    L.mixdir_b200_nccl_unique_id.argtypes = [vp]
    L.mixdir_b200_comm_init.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.mixdir_b200_comm_destroy.argtypes = [vp]
    L.mixdir_b200_create_dist.argtypes = [vp, C.c_int, C.c_int64, C.c_int, ip, vp, C.POINTER(vp)]
    L.mixdir_b200_create_dist_from_r_matrix.argtypes = [vp, C.c_int64, C.c_int, ip, vp, C.POINTER(vp)]
    L.mixdir_b200_destroy.argtypes = [vp]
    L.mixdir_b200_n_rows.restype = C.c_int64
    L.mixdir_b200_n_rows.argtypes = [vp]
    L.mixdir_b200_max_code.argtypes = [vp]
    L.mixdir_b200_n_missing.restype = C.c_int64
    L.mixdir_b200_n_missing.argtypes = [vp]
    L.mixdir_b200_fit.argtypes = [vp, C.POINTER(FitParams), vp, vp, vp, ip, ip, dp, vp, vp, vp, vp,
                                  vp, vp, ip]
    L.mixdir_b200_fit_batch.argtypes = [vp, C.POINTER(FitParams), C.c_int, vp, vp, vp, vp, vp, vp, vp,
                                        vp, vp, vp, vp, vp, ip, vp, vp]
    L.mixdir_b200_predict.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.mixdir_b200_feature_divergence.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, vp, C.c_int64, vp]
    L.mixdir_b200_estep.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp, dp, ip, vp]
    L.mixdir_b200_last_timing.argtypes = [vp, C.POINTER(Timing)]
    L.mixdir_b200_set_engine.argtypes = [vp, C.c_int]
    L.mixdir_b200_set_pairing.argtypes = [vp, C.c_int]
    L.mixdir_b200_unit_plan.argtypes = [C.c_int, ip, C.c_int, vp, ip, ip, vp, vp, vp, vp]
    L.mixdir_b200_unit_layout.argtypes = [C.c_int, ip, C.c_int, vp, C.c_int, ip, ip, ip, vp, vp, vp, vp]
    L.mixdir_b200_last_error.restype = C.c_char_p
    L.mixdir_b200_abi_version.restype = C.c_int
    _lib = L
    return L


def last_error() -> str:
    return lib().mixdir_b200_last_error().decode("utf-8", "replace")
