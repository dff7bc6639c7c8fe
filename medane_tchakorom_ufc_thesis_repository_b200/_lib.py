"""ctypes loader of libmsplit.so (the C-ABI declared in include/msplit.h).

The CUDA library is the product: there is no CPU fallback.  Importing this module never touches
the GPU; ``lib()`` raises ``MsplitError`` loudly when the shared library has not been built.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmsplit.so")

MAX_BLOCKS = 64


class MsplitError(RuntimeError):
    pass


class KspOpts(C.Structure):
    _fields_ = [
        ("restart", C.c_int), ("max_it", C.c_int), ("rtol", C.c_double), ("abstol", C.c_double),
        ("divtol", C.c_double), ("initial_rtol", C.c_int), ("guess_nonzero", C.c_int),
        ("cgs_refine", C.c_int), ("mgs", C.c_int), ("min_it", C.c_int),
    ]


class Problem(C.Structure):
    _fields_ = [
        ("dim", C.c_int), ("m", C.c_int), ("n", C.c_int), ("p", C.c_int), ("block", C.c_int),
        ("nblocks", C.c_int), ("s", C.c_int), ("max_restart", C.c_int), ("keep_csr", C.c_int), ("npb", C.c_int),
    ]


class SolveOpts(C.Structure):
    _fields_ = [
        ("alg", C.c_int), ("s", C.c_int), ("rtol", C.c_double), ("inner", KspOpts), ("max_outer", C.c_int),
        ("record_history", C.c_int), ("profile", C.c_int), ("outer_type", C.c_int), ("outer_max_it", C.c_int),
        ("outer_rtol", C.c_double), ("outer_abstol", C.c_double), ("period", C.c_int * MAX_BLOCKS),
        ("max_seconds", C.c_double), ("detector", C.c_int), ("min_convergence_count", C.c_int), ("max_traversal_ms", C.c_double),
    ]


class Result(C.Structure):
    _fields_ = [
        ("outer_its", C.c_int), ("inner_its_total", C.c_int64), ("norm0", C.c_double), ("last_norm", C.c_double),
        ("final_residual", C.c_double), ("error", C.c_double), ("elapsed_s", C.c_double),
        ("gmres_its", C.c_int), ("gmres_reason", C.c_int), ("gmres_rnorm", C.c_double),
        ("hist_len", C.c_int), ("hist", C.c_double * 4096), ("kernel_launches", C.c_int64),
        ("t_spmv_ms", C.c_double), ("t_mdot_ms", C.c_double), ("t_maxpy_ms", C.c_double), ("t_other_ms", C.c_double),
        ("b_spmv", C.c_double), ("b_mdot", C.c_double), ("b_maxpy", C.c_double), ("b_other", C.c_double),
        ("n_spmv", C.c_int64), ("n_mdot", C.c_int64), ("n_maxpy", C.c_int64), ("n_other", C.c_int64),
        ("stage_inner_s", C.c_double), ("stage_outer_s", C.c_double), ("outer_solver_its", C.c_int64),
        ("hist_dropped", C.c_int), ("stop_reason", C.c_int),
    ]

    def as_dict(self):
        return {
            "outer_its": self.outer_its, "inner_its_total": self.inner_its_total, "norm0": self.norm0,
            "last_norm": self.last_norm, "final_residual": self.final_residual, "error": self.error,
            "elapsed_s": self.elapsed_s, "gmres_its": self.gmres_its, "gmres_reason": self.gmres_reason,
            "gmres_rnorm": self.gmres_rnorm, "hist": np.array(self.hist[: self.hist_len]),
            "kernel_launches": self.kernel_launches, "stage_inner_s": self.stage_inner_s,
            "stage_outer_s": self.stage_outer_s, "outer_solver_its": self.outer_solver_its,
            "hist_dropped": self.hist_dropped, "stop_reason": self.stop_reason,
            "prof": {c: {"ms": getattr(self, f"t_{c}_ms"), "bytes": getattr(self, f"b_{c}"), "launches": getattr(self, f"n_{c}")}
                     for c in ("spmv", "mdot", "maxpy", "other")},
        }


# every symbol include/msplit.h declares (checked by the CPU test-suite)
EXPORTS = [
    "msp_version", "msp_last_error", "msp_device_count", "msp_poisson2d_nnz", "msp_poisson3d_nnz",
    "msp_assemble_poisson2d", "msp_assemble_poisson2d_complete", "msp_assemble_poisson3d", "msp_dimension_related",
    "msp_create", "msp_destroy", "msp_rows", "msp_halo_size", "msp_spmv_format", "msp_persistent_cycles", "msp_mat_nnz", "msp_get_csr", "msp_set_b", "msp_get_b",
    "msp_set_x", "msp_get_x", "msp_set_halo", "msp_get_halo", "msp_get_rhs", "msp_update_local_rhs", "msp_inner_solve",
    "msp_local_residual_norm", "msp_block_residual_norm", "msp_error_norm_sq", "msp_push_iterate", "msp_spmm_AS",
    "msp_minimize_local_qr", "msp_apply_alpha", "msp_tsqr_combine", "msp_op_spmv", "msp_op_mdot", "msp_op_maxpy",
    "msp_bench_kernel", "msp_gmres_solve", "msp_group_create", "msp_group_destroy", "msp_group_engine",
    "msp_group_solve", "msp_comm_unique_id", "msp_comm_init", "msp_comm_export", "msp_comm_connect", "msp_comm_connect_block", "msp_solve",
    "msp_conv_detect_step", "msp_get_solution", "msp_split_blocks", "msp_compute_rhs_ones", "msp_residual_norm",
    "msp_connect_local", "msp_exchange_sync", "msp_async_reset", "msp_exchange_async_publish", "msp_exchange_async_poll",
    "msp_minimize", "msp_set_b_async", "msp_set_x_async", "msp_get_x_async", "msp_copies_wait",
]

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MsplitError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C medane_tchakorom_ufc_thesis_repository_b200/csrc). "
            "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    vp = C.c_void_p
    L.msp_last_error.restype = C.c_char_p
    L.msp_poisson2d_nnz.restype = C.c_int64
    L.msp_poisson2d_nnz.argtypes = [C.c_int] * 4
    L.msp_poisson3d_nnz.restype = C.c_int64
    L.msp_poisson3d_nnz.argtypes = [C.c_int] * 5
    L.msp_assemble_poisson2d.argtypes = [C.c_int] * 5 + [i32p, i32p, f64p]
    L.msp_assemble_poisson2d_complete.argtypes = [C.c_int] * 3 + [i32p, i32p, f64p]
    L.msp_assemble_poisson3d.argtypes = [C.c_int] * 6 + [i32p, i32p, f64p]
    L.msp_dimension_related.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_int)] * 5
    L.msp_create.argtypes = [C.POINTER(Problem), C.c_int, C.POINTER(vp)]
    L.msp_destroy.argtypes = [vp]
    L.msp_rows.argtypes = [vp]
    L.msp_halo_size.argtypes = [vp]
    L.msp_spmv_format.argtypes = [vp, C.POINTER(C.c_int)]
    L.msp_persistent_cycles.argtypes = [vp]
    L.msp_mat_nnz.restype = C.c_int64
    L.msp_mat_nnz.argtypes = [vp, C.c_int]
    L.msp_get_csr.argtypes = [vp, C.c_int, i32p, i32p, f64p]
    for nm in ("msp_set_b", "msp_get_b", "msp_set_x", "msp_get_x", "msp_get_rhs"):
        getattr(L, nm).argtypes = [vp, f64p]
    L.msp_set_halo.argtypes = [vp, C.c_int, f64p]
    L.msp_get_halo.argtypes = [vp, C.c_int, f64p]
    L.msp_update_local_rhs.argtypes = [vp]
    L.msp_inner_solve.argtypes = [vp, C.POINTER(KspOpts), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.msp_local_residual_norm.argtypes = [vp, C.POINTER(C.c_double)]
    L.msp_block_residual_norm.argtypes = [vp, C.POINTER(C.c_double)]
    L.msp_error_norm_sq.argtypes = [vp, C.POINTER(C.c_double)]
    L.msp_push_iterate.argtypes = [vp, C.c_int]
    L.msp_spmm_AS.argtypes = [vp, C.c_int]
    L.msp_minimize_local_qr.argtypes = [vp, C.c_int, f64p]
    L.msp_apply_alpha.argtypes = [vp, C.c_int, f64p]
    L.msp_tsqr_combine.argtypes = [C.c_int, C.c_int, f64p, f64p, C.POINTER(C.c_double)]
    L.msp_op_spmv.argtypes = [vp, C.c_int, f64p, vp, vp, f64p]
    L.msp_op_mdot.argtypes = [vp, C.c_int, f64p, f64p, f64p]
    L.msp_op_maxpy.argtypes = [vp, C.c_int, f64p, f64p, f64p, C.POINTER(C.c_double)]
    L.msp_bench_kernel.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.msp_gmres_solve.argtypes = [vp, C.POINTER(KspOpts), C.POINTER(Result)]
    L.msp_group_create.argtypes = [C.POINTER(Problem), C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]
    L.msp_group_destroy.argtypes = [vp]
    L.msp_group_engine.restype = vp
    L.msp_group_engine.argtypes = [vp, C.c_int]
    L.msp_group_solve.argtypes = [vp, C.POINTER(SolveOpts), C.POINTER(Result)]
    L.msp_comm_unique_id.argtypes = [C.c_char_p]
    L.msp_comm_init.argtypes = [vp, C.c_char_p, C.c_int, C.c_int]
    L.msp_comm_export.argtypes = [vp, C.c_char_p]
    L.msp_comm_connect.argtypes = [vp, C.c_int, C.c_char_p]
    L.msp_comm_connect_block.argtypes = [vp, C.c_int, C.c_char_p]
    L.msp_solve.argtypes = [vp, C.POINTER(SolveOpts), C.POINTER(Result)]
    L.msp_conv_detect_step.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    for nm in ("msp_set_b_async", "msp_set_x_async", "msp_get_x_async"):
        getattr(L, nm).argtypes = [vp, f64p]
    L.msp_copies_wait.argtypes = [vp]
    L.msp_get_solution.argtypes = [vp, f64p]
    L.msp_split_blocks.argtypes = [vp, C.c_int, i32p, i32p, f64p]
    L.msp_compute_rhs_ones.argtypes = [vp]
    L.msp_residual_norm.argtypes = [vp, C.POINTER(C.c_double)]
    L.msp_connect_local.argtypes = [vp, C.c_int, vp]
    L.msp_exchange_sync.argtypes = [vp]
    L.msp_async_reset.argtypes = [vp]
    L.msp_exchange_async_publish.argtypes = [vp, C.c_int]
    L.msp_exchange_async_poll.argtypes = [vp, C.POINTER(C.c_int)]
    L.msp_minimize.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, f64p, C.POINTER(C.c_double)]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc:
        msg = lib().msp_last_error()
        raise MsplitError(msg.decode() if msg else f"libmsplit error {rc}")
