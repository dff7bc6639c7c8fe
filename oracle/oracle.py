"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.  The product package never does.
See oracle/msplit_oracle.h for the parity status (what is pinned by the reference's own
golden vectors and what is "parity unpinned").
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmsplit_oracle.so")

ALG = {
    "SM": 0, "MSM": 0,
    "SMSM_GLOBAL": 1, "SMSM_SEMI_LOCAL": 2, "SMSM_LOCAL": 3,
    "GMRES": 4,
    "AM": 5, "AMAM_GLOBAL": 6, "AMAM_SEMI_LOCAL": 7, "AMAM_LOCAL": 8,
}
OUTER_LSQR, OUTER_QR = 0, 1


class KspOpts(C.Structure):
    _fields_ = [
        ("restart", C.c_int), ("max_it", C.c_int), ("rtol", C.c_double), ("abstol", C.c_double),
        ("divtol", C.c_double), ("initial_rtol", C.c_int), ("guess_nonzero", C.c_int),
        ("cgs_refine", C.c_int), ("mgs", C.c_int), ("min_it", C.c_int),
    ]


class OuterOpts(C.Structure):
    _fields_ = [
        ("type", C.c_int), ("max_it", C.c_int), ("rtol", C.c_double), ("abstol", C.c_double),
        ("lsqr_default_test", C.c_int),
    ]


class Config(C.Structure):
    _fields_ = [
        ("alg", C.c_int), ("dim", C.c_int), ("m", C.c_int), ("n", C.c_int), ("p", C.c_int),
        ("nblocks", C.c_int), ("s", C.c_int), ("rtol", C.c_double),
        ("inner", KspOpts), ("outer", OuterOpts), ("max_outer", C.c_int),
        ("period", C.c_int * 16), ("nthreads", C.c_int), ("max_seconds", C.c_double),
    ]


class Result(C.Structure):
    _fields_ = [
        ("outer_its", C.c_int), ("outer_its_block", C.c_int * 16), ("inner_its_total", C.c_int64),
        ("norm0", C.c_double), ("last_norm", C.c_double), ("final_residual", C.c_double),
        ("error", C.c_double), ("hist_len", C.c_int), ("hist", C.c_double * 4096),
        ("gmres_its", C.c_int), ("gmres_reason", C.c_int), ("gmres_rnorm", C.c_double), ("elapsed_s", C.c_double),
        ("t_outer", C.c_double * 256),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/msplit_oracle.c (gcc, OpenMP) -> oracle/libmsplit_oracle.so."""
    src = os.path.join(_HERE, "msplit_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       env={k: v for k, v in os.environ.items() if k not in ("CC", "CFLAGS")})
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
        f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        L.orc_poisson2d_nnz.restype = C.c_int64
        L.orc_poisson2d_nnz.argtypes = [C.c_int] * 4
        L.orc_poisson2d.argtypes = [C.c_int] * 4 + [i32p, i32p, f64p]
        L.orc_poisson2d_complete.argtypes = [C.c_int] * 2 + [i32p, i32p, f64p]
        L.orc_poisson3d_nnz.restype = C.c_int64
        L.orc_poisson3d_nnz.argtypes = [C.c_int] * 5
        L.orc_poisson3d.argtypes = [C.c_int] * 5 + [i32p, i32p, f64p]
        L.orc_submatrix_nnz.restype = C.c_int64
        L.orc_submatrix_nnz.argtypes = [C.c_int, i32p, i32p, C.c_int, C.c_int]
        L.orc_submatrix.argtypes = [C.c_int, i32p, i32p, f64p, C.c_int, C.c_int, i32p, i32p, f64p]
        L.orc_dimension_related.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_int)] * 5
        L.orc_spmv.restype = None
        L.orc_spmv.argtypes = [C.c_int, i32p, i32p, f64p, f64p, f64p]
        L.orc_residual.restype = None
        L.orc_residual.argtypes = [C.c_int, i32p, i32p, f64p, f64p, f64p, f64p]
        L.orc_dot.restype = C.c_double
        L.orc_dot.argtypes = [C.c_int64, f64p, f64p]
        L.orc_norm2.restype = C.c_double
        L.orc_norm2.argtypes = [C.c_int64, f64p]
        L.orc_block_residual_norm.restype = C.c_double
        L.orc_block_residual_norm.argtypes = [C.c_int, i32p, i32p, f64p, f64p, f64p]
        L.orc_ksp_defaults.restype = None
        L.orc_ksp_defaults.argtypes = [C.POINTER(KspOpts)]
        L.orc_gmres.argtypes = [C.c_int, i32p, i32p, f64p, f64p, f64p, C.POINTER(KspOpts),
                                C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p, C.c_int]
        L.orc_lsqr.argtypes = [C.c_int64, C.c_int, f64p, C.c_int64, f64p, f64p, C.POINTER(OuterOpts),
                               C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_lstsq_qr.argtypes = [C.c_int64, C.c_int, f64p, C.c_int64, f64p, f64p, C.POINTER(C.c_double)]
        L.orc_solve.argtypes = [C.POINTER(Config), C.POINTER(Result), C.c_void_p]
        L.orc_cd_create.restype = C.c_void_p
        L.orc_cd_create.argtypes = [C.c_int]
        L.orc_cd_destroy.restype = None
        L.orc_cd_destroy.argtypes = [C.c_void_p]
        L.orc_cd_step.restype = None
        L.orc_cd_step.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_cd_data_arrival.argtypes = [C.c_void_p] + [C.c_int] * 4
        L.orc_cd_state.argtypes = [C.c_void_p, C.c_int]
        L.orc_cd_phase_tag.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


# ---------------------------------------------------------------- assembly
def poisson2d(m, n, block=0, nblocks=1):
    L = lib()
    nb = (m * n) // nblocks
    nnz = L.orc_poisson2d_nnz(m, n, block, nblocks)
    rp = np.empty(nb + 1, np.int32); ci = np.empty(nnz, np.int32); va = np.empty(nnz, np.float64)
    rc = L.orc_poisson2d(m, n, block, nblocks, rp, ci, va)
    assert rc == 0
    return rp, ci, va


def poisson2d_complete(m, n):
    L = lib()
    nnz = L.orc_poisson2d_nnz(m, n, 0, 1)
    rp = np.empty(m * n + 1, np.int32); ci = np.empty(nnz, np.int32); va = np.empty(nnz, np.float64)
    rc = L.orc_poisson2d_complete(m, n, rp, ci, va)
    if rc:
        raise ValueError("poisson2DMatrix_complete assumes a square mesh (utils.c:390)")
    return rp, ci, va


def poisson3d(nx, ny, nz, block=0, nblocks=1):
    L = lib()
    nb = (nx * ny * nz) // nblocks
    nnz = L.orc_poisson3d_nnz(nx, ny, nz, block, nblocks)
    rp = np.empty(nb + 1, np.int32); ci = np.empty(nnz, np.int32); va = np.empty(nnz, np.float64)
    rc = L.orc_poisson3d(nx, ny, nz, block, nblocks, rp, ci, va)
    assert rc == 0
    return rp, ci, va


def submatrix(rp, ci, va, col_lo, col_hi):
    L = lib()
    nrows = len(rp) - 1
    nnz = L.orc_submatrix_nnz(nrows, rp, ci, col_lo, col_hi)
    orp = np.empty(nrows + 1, np.int32); oci = np.empty(max(nnz, 1), np.int32); ova = np.empty(max(nnz, 1), np.float64)
    L.orc_submatrix(nrows, rp, ci, va, col_lo, col_hi, orp, oci, ova)
    return orp, oci[:nnz], ova[:nnz]


def dimension_related(nprocs, npb, rank, m, n):
    out = [C.c_int() for _ in range(5)]
    rc = lib().orc_dimension_related(nprocs, npb, rank, m, n, *[C.byref(o) for o in out])
    assert rc == 0
    keys = ("njacobi_blocks", "rank_jacobi_block", "proc_local_rank", "n_mesh_points", "jacobi_block_size")
    return dict(zip(keys, (o.value for o in out)))


# ---------------------------------------------------------------- kernels
def spmv(rp, ci, va, x):
    y = np.empty(len(rp) - 1, np.float64)
    lib().orc_spmv(len(rp) - 1, rp, ci, va, np.ascontiguousarray(x, np.float64), y)
    return y


def residual(rp, ci, va, b, x):
    r = np.empty(len(rp) - 1, np.float64)
    lib().orc_residual(len(rp) - 1, rp, ci, va, np.ascontiguousarray(b, np.float64), np.ascontiguousarray(x, np.float64), r)
    return r


def block_residual_norm(rp, ci, va, b, x):
    return lib().orc_block_residual_norm(len(rp) - 1, rp, ci, va, np.ascontiguousarray(b, np.float64),
                                         np.ascontiguousarray(x, np.float64))


def ksp_opts(**kw) -> KspOpts:
    o = KspOpts()
    lib().orc_ksp_defaults(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def gmres(rp, ci, va, b, x0=None, hist_cap=0, **kw):
    o = ksp_opts(**kw)
    n = len(rp) - 1
    x = np.zeros(n) if x0 is None else np.array(x0, np.float64)
    its, reason, rnorm = C.c_int(), C.c_int(), C.c_double()
    hist = np.zeros(max(hist_cap, 1))
    lib().orc_gmres(n, rp, ci, va, np.ascontiguousarray(b, np.float64), x, C.byref(o), C.byref(its), C.byref(reason),
                    C.byref(rnorm), hist.ctypes.data if hist_cap else None, hist_cap)
    return x, its.value, reason.value, rnorm.value


def lsqr(R, b, max_it=100, rtol=1e-15, abstol=1e-100, default_test=True):
    """R: (nrows, s) array (any layout); solved column-major like a PETSc MATDENSE."""
    Rf = np.asfortranarray(R, np.float64)
    nrows, s = Rf.shape
    o = OuterOpts(OUTER_LSQR, max_it, rtol, abstol, int(default_test))
    alpha = np.zeros(s)
    its, reason, rnorm = C.c_int(), C.c_int(), C.c_double()
    flat = np.ascontiguousarray(Rf.T).reshape(-1)
    lib().orc_lsqr(nrows, s, flat, nrows, np.ascontiguousarray(b, np.float64), alpha, C.byref(o), C.byref(its),
                   C.byref(reason), C.byref(rnorm))
    return alpha, its.value, reason.value, rnorm.value


def lstsq_qr(R, b):
    Rf = np.asfortranarray(R, np.float64)
    nrows, s = Rf.shape
    alpha = np.zeros(s)
    rnorm = C.c_double()
    flat = np.ascontiguousarray(Rf.T).reshape(-1)
    rc = lib().orc_lstsq_qr(nrows, s, flat, nrows, np.ascontiguousarray(b, np.float64), alpha, C.byref(rnorm))
    if rc:
        raise np.linalg.LinAlgError(f"orc_lstsq_qr rc={rc}")
    return alpha, rnorm.value


# ---------------------------------------------------------------- outer loops
def solve(alg, m, n, p=1, nblocks=2, s=4, rtol=1e-6, inner=None, outer_type="qr", outer_max_it=100,
          outer_rtol=1e-15, max_outer=0, periods=None, nthreads=0, want_x=True, max_seconds=0.0):
    """Run one of the reference's drivers on the oracle.  ``inner`` is a dict of KspOpts fields."""
    cfg = Config()
    cfg.alg = ALG[alg] if isinstance(alg, str) else int(alg)
    cfg.dim = 3 if p > 1 else 2
    cfg.m, cfg.n, cfg.p = m, n, p
    cfg.nblocks, cfg.s, cfg.rtol = nblocks, s, rtol
    cfg.inner = ksp_opts(**(inner or {}))
    cfg.outer = OuterOpts(OUTER_LSQR if outer_type == "lsqr" else OUTER_QR, outer_max_it, outer_rtol, 1e-100, 1)
    cfg.max_outer = max_outer
    for i in range(16):
        cfg.period[i] = (periods[i] if periods and i < len(periods) else 1)
    cfg.nthreads = nthreads
    cfg.max_seconds = max_seconds
    res = Result()
    ntot = m * n * max(p, 1)
    x = np.zeros(ntot) if want_x else None
    rc = lib().orc_solve(C.byref(cfg), C.byref(res), x.ctypes.data if want_x else None)
    out = {
        "rc": rc, "outer_its": res.outer_its, "outer_its_block": list(res.outer_its_block)[:nblocks],
        "inner_its_total": res.inner_its_total, "norm0": res.norm0, "last_norm": res.last_norm,
        "final_residual": res.final_residual, "error": res.error,
        "hist": np.array(res.hist[: res.hist_len]),
        "gmres_its": res.gmres_its, "gmres_reason": res.gmres_reason, "gmres_rnorm": res.gmres_rnorm, "x": x,
        "elapsed_s": res.elapsed_s, "t_outer": list(res.t_outer)[: min(res.outer_its, 256)],
    }
    return out


class ConvDetect:
    """conv_detection_prime.c state machines of all block roots, with last-value mailboxes."""

    NORMAL, WAIT4VERIFICATION, VERIFICATION, FINISHED = range(4)

    def __init__(self, nblocks):
        self.h = lib().orc_cd_create(nblocks)
        self.nblocks = nblocks

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_cd_destroy(self.h)
            self.h = None

    def step(self, k, under):
        lib().orc_cd_step(self.h, k, int(bool(under)))

    def data_arrival(self, k, src, tag, it):
        return bool(lib().orc_cd_data_arrival(self.h, k, src, tag, it))

    def state(self, k):
        return lib().orc_cd_state(self.h, k)

    def phase_tag(self, k):
        return lib().orc_cd_phase_tag(self.h, k)
