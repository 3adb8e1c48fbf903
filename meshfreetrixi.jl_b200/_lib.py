"""ctypes binding of libmft_b200.so (the C ABI declared in include/mft_b200.h).

The library is the product; there is no Python/NumPy/CPU fallback.  If the shared object is missing or a call fails
the error is raised, never papered over.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# MFT_LIB_PATH: a prebuilt variant of the same library (kernel-tuning experiments: other launch bounds / tile sizes)
LIB_PATH = os.environ.get("MFT_LIB_PATH") or os.path.join(HERE, "libmft_b200.so")

# constants of include/mft_b200.h
EQ_EULER2D, EQ_ADVECTION2D = 0, 1
OP_DX, OP_DY = 0, 1
BC_DIRICHLET, BC_SLIP_WALL, BC_DO_NOTHING = 0, 1, 2
SRC_HV_FLYER, SRC_HV_TOMINEC, SRC_UPWIND, SRC_RESIDUAL, SRC_IGR = 0, 1, 2, 3, 4
MEM_HOST, MEM_DEVICE = 0, 1
OPT_EXACT_ORDER, OPT_MEAN_DIVISOR_VN, OPT_MAX_LEXICOGRAPHIC, OPT_DIAGNOSTICS, OPT_CUDA_GRAPH = 0, 1, 2, 3, 4
OPT_STAGE_WEIGHTS, OPT_PREFETCH_DISTANCE, OPT_REFINE_ORDER, OPT_SINGLE_SWEEP_EXACT = 5, 6, 7, 8
OPT_PAIR_ROWS = 9
OPT_TILE = 10
OPT_TILE_ROWS = 11
OPT_FUSED_STEP = 12
OPT_PDL = 13
OPT_LAYOUT_DEVICE = 14
VAR_DENSITY, VAR_PRESSURE = 0, 1
FIELD_EPS, FIELD_EPS_UW, FIELD_EPS_RV, FIELD_EPS_C, FIELD_RESIDUAL, FIELD_APPROX_DU, FIELD_NORMS, FIELD_SIGMA, FIELD_IGR_STATUS, FIELD_NORM_MISSES = range(10)
SSPRK33 = 0
K_PASS_A, K_PASS_B, K_REDUCE, K_STAGE, K_BC, K_OTHER = range(6)

EXPORTS = [
    "mft_last_error", "mft_version", "mft_device_count", "mft_ctx_create", "mft_ctx_destroy", "mft_set_equation",
    "mft_set_option", "mft_set_permutation", "mft_set_order_keys", "mft_set_operator_csc", "mft_set_operator_ell",
    "mft_add_boundary", "mft_update_boundary_values", "mft_set_stage_boundary_values", "mft_add_source", "mft_finalize", "mft_rhs", "mft_calc_fluxes",
    "mft_apply_source", "mft_boundary_pass", "mft_upload_state", "mft_download_state", "mft_download_du",
    "mft_history_push", "mft_history_push_weights", "mft_ssprk_step", "mft_ssprk43_step", "mft_step_commit", "mft_get_field", "mft_synchronize", "mft_count_nonfinite",
    "mft_launch_count", "mft_timer_start", "mft_timer_stop", "mft_kernel_time_ms", "mft_set_kernel_timing", "mft_host_alloc", "mft_host_free",
    "mft_host_register", "mft_host_unregister", "mft_sfc_order", "mft_debug_tile_selftest", "mft_debug_tile_selftest_csr", "mft_debug_tile_build_compare", "mft_debug_tile_build_compare_csr", "mft_debug_matcher_compare", "mft_debug_tiler_checksum", "mft_setup_knn", "mft_setup_knn_queries", "mft_setup_rbf_weights", "mft_setup_rbf_weights_rows", "mft_setup_rbf_weights_hybrid", "mft_set_neighbors", "mft_limiter_zhang_shu", "mft_set_stage_limiter", "mft_nccl_unique_id", "mft_comm_init", "mft_set_halo", "mft_p2p_handles", "mft_p2p_connect",
]

_lib = None


class MftError(RuntimeError):
    pass


def load():
    """Load libmft_b200.so; raises if it has not been built (see build.py / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MftError(f"{LIB_PATH} not found: build it with `python meshfreetrixi.jl_b200/build.py` "
                       "(needs nvcc).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.mft_last_error.restype = C.c_char_p
    lib.mft_launch_count.restype = C.c_int64
    lib.mft_launch_count.argtypes = [C.c_void_p]
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    sigs = {
        "mft_ctx_create": [C.POINTER(vp), i32, i64, i64, i32, i32, i32],
        "mft_ctx_destroy": [vp],
        "mft_set_neighbors": [vp, vp],
        "mft_count_nonfinite": [vp, vp],
        "mft_setup_knn": [i32, i64, vp, vp, i32, vp, vp],
        "mft_setup_knn_queries": [i32, i64, vp, vp, i32, i64, vp, vp, vp],
        "mft_setup_rbf_weights_rows": [i32, i64, vp, vp, i64, i32, vp, i32, i32, i32, vp, vp],
        "mft_setup_rbf_weights_hybrid": [i32, i64, vp, vp, i64, i32, vp, i32, dbl, dbl, dbl, i32, i32, vp, vp],
        "mft_setup_rbf_weights": [i32, i64, vp, vp, i32, vp, i32, i32, i32, vp, vp],
        "mft_limiter_zhang_shu": [vp, i32, vp, vp, vp, i32],
        "mft_set_stage_limiter": [vp, i32, vp, vp],
        "mft_debug_tile_selftest": [i64, i32, i32, i32, i32, C.c_uint, vp],
        "mft_debug_tile_selftest_csr": [i64, i64, vp, vp, i32, i32, C.c_uint, vp],
        "mft_debug_tile_build_compare": [i64, i32, i32, i32, C.c_uint],
        "mft_debug_tile_build_compare_csr": [i64, i64, vp, vp, i32, C.c_uint],
        "mft_debug_matcher_compare": [C.c_uint, i32, C.POINTER(C.c_longlong)],
        "mft_debug_tiler_checksum": [vp, i32, C.POINTER(C.c_ulonglong), C.POINTER(C.c_longlong)],
        "mft_set_equation": [vp, i32, vp, i32],
        "mft_set_option": [vp, i32, dbl],
        "mft_set_permutation": [vp, vp],
        "mft_set_order_keys": [vp, vp],
        "mft_set_operator_csc": [vp, i32, vp, vp, vp],
        "mft_set_operator_ell": [vp, vp, vp, vp],
        "mft_add_boundary": [vp, i32, i64, vp, vp, vp],
        "mft_update_boundary_values": [vp, i32, vp],
        "mft_set_stage_boundary_values": [vp, i32, i32, vp],
        "mft_add_source": [vp, i32, vp, i32, vp, vp, vp],
        "mft_finalize": [vp],
        "mft_rhs": [vp, dbl, vp, vp, i32],
        "mft_calc_fluxes": [vp, vp, vp],
        "mft_apply_source": [vp, i32, dbl, vp, vp],
        "mft_boundary_pass": [vp, dbl, vp, vp],
        "mft_upload_state": [vp, vp],
        "mft_download_state": [vp, vp],
        "mft_download_du": [vp, vp],
        "mft_history_push": [vp, dbl, i64, i32],
        "mft_history_push_weights": [vp, dbl, i64, i32, vp],
        "mft_ssprk_step": [vp, i32, dbl, dbl],
        "mft_ssprk43_step": [vp, dbl, dbl, dbl, dbl, C.POINTER(dbl), C.POINTER(i64)],
        "mft_step_commit": [vp, i32],
        "mft_get_field": [vp, i32, vp],
        "mft_synchronize": [vp],
        "mft_kernel_time_ms": [vp, i32, C.POINTER(dbl), C.POINTER(i64)],
        "mft_set_kernel_timing": [vp, i32],
        "mft_timer_start": [vp],
        "mft_timer_stop": [vp, C.POINTER(dbl)],
        "mft_host_alloc": [C.POINTER(vp), i64],
        "mft_host_free": [vp],
        "mft_host_register": [vp, i64],
        "mft_host_unregister": [vp],
        "mft_sfc_order": [i64, vp, vp, vp],
        "mft_nccl_unique_id": [vp],
        "mft_comm_init": [vp, i32, i32, vp],
        "mft_set_halo": [vp, i32, vp, vp, vp, vp],
        "mft_p2p_handles": [vp, vp],
        "mft_p2p_connect": [vp, i32, i32, vp, vp, i64],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise MftError(f"libmft_b200 error {rc}: {load().mft_last_error().decode()}")


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def soa_ptrs(arr: np.ndarray):
    """(V,N) C-contiguous float64 -> array of V row pointers (= StructArrays.components(u))."""
    assert arr.dtype == np.float64 and arr.flags.c_contiguous and arr.ndim == 2
    V = arr.shape[0]
    ptrs = (C.c_void_p * V)(*[arr.ctypes.data + v * arr.strides[0] for v in range(V)])
    return ptrs


def sfc_order(points: np.ndarray) -> np.ndarray:
    """Hilbert ordering (host helper in the library): returns perm (0-based), device row d <- point perm[d]."""
    x = np.ascontiguousarray(points[:, 0], dtype=np.float64)
    y = np.ascontiguousarray(points[:, 1], dtype=np.float64)
    out = np.empty(len(x), dtype=np.int64)
    check(load().mft_sfc_order(len(x), ptr(x), ptr(y), ptr(out)))
    return out - 1


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array backed by cudaHostAlloc memory (never freed explicitly; process-lifetime buffers)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    check(load().mft_host_alloc(C.byref(p), n))
    buf = (C.c_char * n).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)
