"""Host-side mirror of the reference's operator surface for the multisplitting solve path.

Names, argument meaning and error behaviour follow include/utils.h + include/comm.h of the
reference (file:line cited per function); PETSc Mat/Vec/KSP handles are replaced by a per-block
``Engine`` that lives on one GPU.  Everything here forwards to the C-ABI of libmsplit.so
(include/msplit.h); numpy arrays are host buffers only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import KspOpts, MsplitError, Problem, Result, SolveOpts, check

ALG = {
    "SM": 0, "MSM": 0, "SMSM_GLOBAL": 1, "SMSM_SEMI_LOCAL": 2, "SMSM_LOCAL": 3, "GMRES": 4,
    "AM": 5, "AMAM_GLOBAL": 6, "AMAM_SEMI_LOCAL": 7, "AMAM_LOCAL": 8,
}
MAT_STRIP, MAT_DIAG, MAT_OFFDIAG = 0, 1, 2


def ksp_opts(restart=30, max_it=10000, rtol=1e-5, abstol=1e-50, divtol=1e4, initial_rtol=0, guess_nonzero=0,
             cgs_refine=0, mgs=0, min_it=0) -> KspOpts:
    """PETSc 3.22.1 KSP defaults (tmp/petscmpiexec_help:336-342,602-615)."""
    return KspOpts(restart, max_it, rtol, abstol, divtol, initial_rtol, guess_nonzero, cgs_refine, mgs, min_it)


# ------------------------------------------------------------------ assembly (bit-exact gate)
def computeDimensionRelatedVariables(nprocs, nprocs_per_jacobi_block, proc_global_rank, n_mesh_lines, n_mesh_columns):
    """utils.c:652-666."""
    out = [C.c_int() for _ in range(5)]
    check(_lib.lib().msp_dimension_related(nprocs, nprocs_per_jacobi_block, proc_global_rank, n_mesh_lines,
                                           n_mesh_columns, *[C.byref(o) for o in out]))
    keys = ("njacobi_blocks", "rank_jacobi_block", "proc_local_rank", "n_mesh_points", "jacobi_block_size")
    return dict(zip(keys, (o.value for o in out)))


def poisson2DMatrix(n_grid_lines, n_grid_columns, rank_jacobi_block=0, njacobi_blocks=1, device=0):
    """utils.c:247-293 + MatAssembly: CSR (rowptr, colidx[global], val) of the block's strip."""
    L = _lib.lib()
    nb = (n_grid_lines * n_grid_columns) // njacobi_blocks
    nnz = L.msp_poisson2d_nnz(n_grid_lines, n_grid_columns, rank_jacobi_block, njacobi_blocks)
    rp = np.empty(nb + 1, np.int32); ci = np.empty(nnz, np.int32); va = np.empty(nnz, np.float64)
    check(L.msp_assemble_poisson2d(device, n_grid_lines, n_grid_columns, rank_jacobi_block, njacobi_blocks, rp, ci, va))
    return rp, ci, va


def poisson2DMatrix_complete(n_mesh_lines, n_mesh_columns, device=0):
    """utils.c:383-445 (square meshes only, as in the reference)."""
    L = _lib.lib()
    nnz = L.msp_poisson2d_nnz(n_mesh_lines, n_mesh_columns, 0, 1)
    rp = np.empty(n_mesh_lines * n_mesh_columns + 1, np.int32); ci = np.empty(nnz, np.int32); va = np.empty(nnz, np.float64)
    check(L.msp_assemble_poisson2d_complete(device, n_mesh_lines, n_mesh_columns, rp, ci, va))
    return rp, ci, va


def poisson3DMatrix(n_grid_lines, n_grid_columns, n_grid_depth, rank_jacobi_block=0, njacobi_blocks=1, device=0):
    """utils.c:30-121."""
    L = _lib.lib()
    nb = (n_grid_lines * n_grid_columns * n_grid_depth) // njacobi_blocks
    nnz = L.msp_poisson3d_nnz(n_grid_lines, n_grid_columns, n_grid_depth, rank_jacobi_block, njacobi_blocks)
    rp = np.empty(nb + 1, np.int32); ci = np.empty(nnz, np.int32); va = np.empty(nnz, np.float64)
    check(L.msp_assemble_poisson3d(device, n_grid_lines, n_grid_columns, n_grid_depth, rank_jacobi_block,
                                   njacobi_blocks, rp, ci, va))
    return rp, ci, va


# ------------------------------------------------------------------ one Jacobi block on one GPU
class Engine:
    """One Jacobi block resident on one GPU (msp_engine)."""

    def __init__(self, m, n, p=1, block=0, nblocks=1, s=0, max_restart=30, device=0, keep_csr=False, _handle=None,
                 _owner=None, npb=1):
        """block / nblocks: this GPU's strip and the number of strips; npb: GPUs per Jacobi block (Jacobi blocks = nblocks / npb)."""
        self._owner = _owner
        if _handle is not None:
            self.h = _handle
        else:
            prob = Problem(3 if p > 1 else 2, m, n, p, block, nblocks, s, max_restart, int(keep_csr), int(npb))
            h = C.c_void_p()
            check(_lib.lib().msp_create(C.byref(prob), device, C.byref(h)))
            self.h = h
        self.s = s
        self.nb = _lib.lib().msp_rows(self.h)
        self.H = _lib.lib().msp_halo_size(self.h)

    def spmv_format(self):
        """("cdia" | "dia" | "ell", width) of the storage the hot SpMV reads."""
        w = C.c_int()
        kind = _lib.lib().msp_spmv_format(self.h, C.byref(w))
        return {2: "cdia", 1: "dia"}.get(kind, "ell"), w.value

    def persistent_cycles(self):
        """True when each GMRES restart cycle of this block runs as one persistent cooperative kernel (small blocks)."""
        return _lib.lib().msp_persistent_cycles(self.h) == 1

    @staticmethod
    def spmv_bytes_per_row(fmt, width):
        """Bytes per row the SpMV of that storage streams (matrix + x once + y once)."""
        return {"cdia": 1.0, "dia": 8.0 * width, "ell": 12.0 * width}[fmt] + 16.0

    def close(self):
        if self._owner is None and getattr(self, "h", None):
            _lib.lib().msp_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # divideSubDomainIntoBlockMatrices utils.c:450-478
    def divideSubDomainIntoBlockMatrices(self, which):
        L = _lib.lib()
        nnz = L.msp_mat_nnz(self.h, which)
        if nnz < 0:
            check(1)
        rp = np.empty(self.nb + 1, np.int32); ci = np.empty(max(nnz, 1), np.int32); va = np.empty(max(nnz, 1), np.float64)
        check(L.msp_get_csr(self.h, which, rp, ci, va))
        return rp, ci[:nnz], va[:nnz]

    def _getv(self, fn, n):
        out = np.empty(n, np.float64)
        check(fn(self.h, out))
        return out

    @property
    def b(self):
        return self._getv(_lib.lib().msp_get_b, self.nb)

    @b.setter
    def b(self, v):
        check(_lib.lib().msp_set_b(self.h, np.ascontiguousarray(v, np.float64)))

    @property
    def x(self):
        return self._getv(_lib.lib().msp_get_x, self.nb)

    @x.setter
    def x(self, v):
        check(_lib.lib().msp_set_x(self.h, np.ascontiguousarray(v, np.float64)))

    def get_x(self, out=None):
        """msp_get_x straight into a caller-owned (e.g. pinned) host array."""
        if out is None:
            return self.x
        check(_lib.lib().msp_get_x(self.h, out))
        return out

    # pipelined transfers from / to page-locked host arrays (valid until copies_wait)
    def set_b_async(self, v):
        check(_lib.lib().msp_set_b_async(self.h, v))

    def set_x_async(self, v):
        check(_lib.lib().msp_set_x_async(self.h, v))

    def get_x_async(self, out):
        check(_lib.lib().msp_get_x_async(self.h, out))

    def copies_wait(self):
        check(_lib.lib().msp_copies_wait(self.h))

    @property
    def rhs(self):
        return self._getv(_lib.lib().msp_get_rhs, self.nb)

    def set_halo(self, side, v):
        check(_lib.lib().msp_set_halo(self.h, side, np.ascontiguousarray(v, np.float64)))

    def get_halo(self, side):
        out = np.empty(self.H, np.float64)
        check(_lib.lib().msp_get_halo(self.h, side, out))
        return out

    def updateLocalRHS(self):
        """utils.c:943-948: rhs_K = b_K - A_KJ x_J."""
        check(_lib.lib().msp_update_local_rhs(self.h))

    def inner_solver(self, opts: KspOpts):
        """utils.c:950-970: UIR norm, nonzero guess, KSPSolve(A_KK, rhs_K, x_K); returns (its, reason, rnorm)."""
        its, reason, rn = C.c_int(), C.c_int(), C.c_double()
        check(_lib.lib().msp_inner_solve(self.h, C.byref(opts), C.byref(its), C.byref(reason), C.byref(rn)))
        return its.value, reason.value, rn.value

    def local_residual_norm(self):
        out = C.c_double()
        check(_lib.lib().msp_local_residual_norm(self.h, C.byref(out)))
        return out.value

    def block_residual_norm(self):
        out = C.c_double()
        check(_lib.lib().msp_block_residual_norm(self.h, C.byref(out)))
        return out.value

    def push_iterate(self, t):
        check(_lib.lib().msp_push_iterate(self.h, t))

    def spmm_AS(self, kind):
        check(_lib.lib().msp_spmm_AS(self.h, ALG[kind] if isinstance(kind, str) else kind))

    def minimize_local_qr(self, kind):
        u = np.zeros((self.s + 1) * (self.s + 1))
        check(_lib.lib().msp_minimize_local_qr(self.h, ALG[kind] if isinstance(kind, str) else kind, u))
        return u

    def apply_alpha(self, kind, alpha):
        check(_lib.lib().msp_apply_alpha(self.h, ALG[kind] if isinstance(kind, str) else kind,
                                         np.ascontiguousarray(alpha, np.float64)))

    # ---- the exchange and the minimiser as separate calls (include/comm.h, include/utils.h of the reference)
    def connect_local(self, side, neighbour: "Engine"):
        """Wire a neighbour engine of the same process (side 0: block K-1, side 1: block K+1)."""
        check(_lib.lib().msp_connect_local(self.h, side, neighbour.h))

    def compute_rhs_ones(self):
        """computeTheRightHandSideWithInitialGuess utils.c:623-650: b_K = A_K,: 1, rhs_K = b_K, halos zero."""
        check(_lib.lib().msp_compute_rhs_ones(self.h))

    def comm_sync_send_and_receive(self):
        """comm.c:126-141 (collective: one caller per block)."""
        check(_lib.lib().msp_exchange_sync(self.h))

    def async_reset(self):
        check(_lib.lib().msp_async_reset(self.h))

    def comm_async_test_and_send(self, iteration):
        """comm.c:531-554."""
        check(_lib.lib().msp_exchange_async_publish(self.h, iteration))

    def comm_async_probe_and_receive(self):
        """comm.c:455-529; returns [accepted_lower, accepted_upper]."""
        acc = (C.c_int * 2)()
        check(_lib.lib().msp_exchange_async_poll(self.h, acc))
        return [int(acc[0]), int(acc[1])]

    def outer_solver_norm_equation(self, kind, outer_type="tsqr", outer_max_it=100, outer_rtol=1e-15):
        """utils.c:1061-1103 (+ the outer-solver menu utils.c:972-1043): returns (alpha, ||rhs - R alpha||); x = S alpha is applied."""
        alpha = np.zeros(self.s)
        rn = C.c_double()
        ot = {"tsqr": 0, "qr": 0, "lsqr": 1, "gram": 2, "normal": 2, "cg": 3, "cgne": 4}[outer_type]
        check(_lib.lib().msp_minimize(self.h, ALG[kind] if isinstance(kind, str) else kind, ot, outer_max_it, outer_rtol, alpha, C.byref(rn)))
        return alpha, rn.value

    def computeFinalResidualNorm(self):
        """utils.c:575-595 (collective)."""
        out = C.c_double()
        check(_lib.lib().msp_residual_norm(self.h, C.byref(out)))
        return out.value

    def get_solution(self):
        out = np.empty(self.nb, np.float64)
        check(_lib.lib().msp_get_solution(self.h, out))
        return out

    # raw kernels on host data
    def spmv(self, which, x, halo_lo=None, halo_hi=None):
        y = np.empty(self.nb)
        lo = np.ascontiguousarray(halo_lo, np.float64) if halo_lo is not None else None
        hi = np.ascontiguousarray(halo_hi, np.float64) if halo_hi is not None else None
        check(_lib.lib().msp_op_spmv(self.h, which, np.ascontiguousarray(x, np.float64),
                                     lo.ctypes.data if lo is not None else None,
                                     hi.ctypes.data if hi is not None else None, y))
        return y

    def mdot(self, V, w):
        V = np.ascontiguousarray(V, np.float64)
        h = np.empty(V.shape[0])
        check(_lib.lib().msp_op_mdot(self.h, V.shape[0], V.reshape(-1), np.ascontiguousarray(w, np.float64), h))
        return h

    def maxpy(self, V, coef, w):
        V = np.ascontiguousarray(V, np.float64)
        w = np.array(w, np.float64)
        nrm = C.c_double()
        check(_lib.lib().msp_op_maxpy(self.h, V.shape[0], V.reshape(-1), np.ascontiguousarray(coef, np.float64), w, C.byref(nrm)))
        return w, nrm.value

    def bench_kernel(self, op, nv=0, iters=20, flush_l2=False):
        ms = C.c_double()
        check(_lib.lib().msp_bench_kernel(self.h, op, nv, iters, int(flush_l2), C.byref(ms)))
        return ms.value

    def gmres_solve(self, opts: KspOpts):
        """gmres_solution.c:50-85."""
        res = Result()
        check(_lib.lib().msp_gmres_solve(self.h, C.byref(opts), C.byref(res)))
        return res.as_dict()

    # one process per GPU plumbing
    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        check(_lib.lib().msp_comm_init(self.h, unique_id, rank, nranks))

    def comm_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(_lib.lib().msp_comm_export(self.h, buf))
        return buf.raw

    def comm_connect(self, side: int, handle: bytes):
        check(_lib.lib().msp_comm_connect(self.h, side, handle))

    def comm_connect_block(self, block: int, handle: bytes):
        check(_lib.lib().msp_comm_connect_block(self.h, block, handle))

    def solve(self, alg, s=0, rtol=1e-6, inner: Optional[KspOpts] = None, max_outer=0, record_history=True, profile=False,
              **outer):
        o = make_solve_opts(alg, s, rtol, inner, max_outer, record_history, profile=profile, **outer)
        res = Result()
        check(_lib.lib().msp_solve(self.h, C.byref(o), C.byref(res)))
        return res.as_dict()


def tsqr_combine(s, factors):
    """Root of the TSQR tree: stacked (s+1)x(s+1) factors -> (alpha, ||b - R alpha||)."""
    f = np.ascontiguousarray(np.concatenate([np.asarray(u, np.float64).reshape(-1) for u in factors]))
    alpha = np.zeros(s)
    rn = C.c_double()
    check(_lib.lib().msp_tsqr_combine(s, len(factors), f, alpha, C.byref(rn)))
    return alpha, rn.value


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(_lib.lib().msp_comm_unique_id(buf))
    return buf.raw


def make_solve_opts(alg, s=0, rtol=1e-6, inner: Optional[KspOpts] = None, max_outer=0, record_history=True,
                    periods: Optional[Sequence[int]] = None, profile=False, outer_type="tsqr", outer_max_it=100,
                    outer_rtol=1e-15, outer_abstol=1e-100, max_seconds=0.0, detector="prime", min_convergence_count=4,
                    max_traversal_ms=0.0) -> SolveOpts:
    o = SolveOpts()
    o.alg = ALG[alg] if isinstance(alg, str) else int(alg)
    o.s = s
    o.rtol = rtol
    o.inner = inner if inner is not None else ksp_opts()
    o.max_outer = max_outer
    o.record_history = int(record_history)
    o.profile = int(profile)
    o.outer_type = {"tsqr": 0, "qr": 0, "lsqr": 1, "gram": 2, "normal": 2, "cg": 3, "cgne": 4}[outer_type]
    o.outer_max_it, o.outer_rtol, o.outer_abstol = outer_max_it, outer_rtol, outer_abstol
    o.max_seconds = float(max_seconds)
    o.detector = {"prime": 0, "legacy": 1}[detector]
    o.min_convergence_count = int(min_convergence_count)
    o.max_traversal_ms = float(max_traversal_ms)
    for i in range(_lib.MAX_BLOCKS):
        o.period[i] = periods[i] if periods and i < len(periods) else 0
    return o


class Group:
    """All Jacobi blocks in one process, one host thread per block (msp_group).  Blocks may share a GPU
    (tests) or sit on different GPUs of one box (peer access)."""

    def __init__(self, m, n, p=1, nblocks=2, s=0, max_restart=30, devices: Optional[Sequence[int]] = None,
                 keep_csr=False, npb=1):
        """nblocks Jacobi blocks of npb GPUs each: nblocks * npb engines (strips), in strip order."""
        nranks = nblocks * npb
        prob = Problem(3 if p > 1 else 2, m, n, p, 0, nranks, s, max_restart, int(keep_csr), int(npb))
        devs = (C.c_int * nranks)(*(devices if devices is not None else [0] * nranks))
        h = C.c_void_p()
        check(_lib.lib().msp_group_create(C.byref(prob), nranks, devs, C.byref(h)))
        self.h = h
        self.nblocks = nranks   # engines; Jacobi blocks = nranks / npb
        self.njacobi_blocks = nblocks
        self.npb = npb
        self.s = s
        self.engines = [Engine(m, n, p, k, nranks, s, max_restart, _handle=C.c_void_p(_lib.lib().msp_group_engine(h, k)),
                               _owner=self) for k in range(nranks)]

    def close(self):
        if getattr(self, "h", None):
            _lib.lib().msp_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, alg, s=0, rtol=1e-6, inner: Optional[KspOpts] = None, max_outer=0, record_history=True, periods=None,
              profile=False, **outer):
        o = make_solve_opts(alg, s, rtol, inner, max_outer, record_history, periods, profile, **outer)
        res = (Result * self.nblocks)()
        check(_lib.lib().msp_group_solve(self.h, C.byref(o), res))
        return [r.as_dict() for r in res]

    def solution(self):
        return np.concatenate([e.x for e in self.engines])
