"""An INDEPENDENT numpy/scipy statement of the reference's solve path (tests only).

Written without looking at oracle/msplit_oracle.c's control flow, from the reference drivers
(src/synchronous-multisplitting/synchronous-multisplitting.c:170-206,
src/synchronous-multisplitting-synchronous-minimization-global/…-global.c:288-363) and the documented PETSc
semantics (KSPGMRES with a nonzero initial guess, the initial-residual-norm convergence test, `max_it`;
tmp/petscmpiexec_help:336-342,602-615).  Deliberately different in every implementation choice:

  * the matrix is scipy.sparse (kron of 1-D stencils), not a hand-assembled CSR;
  * GMRES keeps the Hessenberg matrix and solves the small least-squares problem with numpy.linalg.lstsq
    at every step (no Givens recurrence) and orthogonalises with classical Gram-Schmidt as a dense mat-vec;
  * the minimisation is numpy.linalg.lstsq on R = A S with the raw iterates as columns.

Agreement between this file and the oracle therefore pins the oracle's solver arithmetic from a second side
(VERDICT r01 "third, independent pin").
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def poisson2d(m, n):
    """5-point matrix of utils.c:247-293, row Ii = i*n + j (i: grid line, j: grid column)."""
    def lap(k):
        return sp.diags([-np.ones(k - 1), 2.0 * np.ones(k), -np.ones(k - 1)], [-1, 0, 1])
    return (sp.kron(lap(m), sp.identity(n)) + sp.kron(sp.identity(m), lap(n))).tocsr()


def poisson3d(nx, ny, nz):
    """7-point matrix of utils.c:30-121, row = i + j*nx + k*nx*ny."""
    def lap(k):
        return sp.diags([-np.ones(k - 1), 2.0 * np.ones(k), -np.ones(k - 1)], [-1, 0, 1])
    ix, iy, iz = sp.identity(nx), sp.identity(ny), sp.identity(nz)
    return (sp.kron(iz, sp.kron(iy, lap(nx))) + sp.kron(iz, sp.kron(lap(ny), ix)) + sp.kron(lap(nz), sp.kron(iy, ix))).tocsr()


def gmres_capped(A, b, x, restart=30, max_it=20, rtol=1e-10, abstol=1e-100, refine=0):
    """Restarted GMRES as PETSc runs it inside inner_solver (utils.c:950-970): nonzero initial guess, tolerance relative to
    the INITIAL residual, no convergence verdict before the first iteration, stop after max_it iterations in total
    (the iterate of the truncated cycle is kept).  Returns (x, iterations).

    refine: -ksp_gmres_cgs_refinement_type (tmp/petscmpiexec_help:607) — 0 refine_never, 2 refine_always (a second
    classical Gram-Schmidt pass on the orthogonalised vector, its coefficients added to the Hessenberg column),
    1 refine_ifneeded (the second pass only when the vector lost more than it kept: ||w_after|| < ||h||)."""
    its = 0
    r0norm = None
    while True:
        r = b - A @ x
        beta = np.linalg.norm(r)
        if r0norm is None:
            r0norm = beta
        tol = max(rtol * r0norm, abstol)
        if beta == 0.0 or (its > 0 and beta <= tol) or its >= max_it:
            return x, its
        k = min(restart, max_it - its)
        V = np.zeros((k + 1, b.size))
        H = np.zeros((k + 1, k))
        V[0] = r / beta
        y = None
        used = 0
        for j in range(k):
            w = A @ V[j]
            h = V[: j + 1] @ w            # classical Gram-Schmidt, first pass
            w = w - h @ V[: j + 1]
            if refine == 2 or (refine == 1 and np.linalg.norm(w) < np.linalg.norm(h)):
                h2 = V[: j + 1] @ w       # second pass on the already orthogonalised vector
                w = w - h2 @ V[: j + 1]
                h = h + h2
            H[: j + 1, j] = h
            H[j + 1, j] = np.linalg.norm(w)
            used = j + 1
            e1 = np.zeros(j + 2)
            e1[0] = beta
            y, *_ = np.linalg.lstsq(H[: j + 2, : j + 1], e1, rcond=None)
            res = np.linalg.norm(e1 - H[: j + 2, : j + 1] @ y)
            its += 1
            if H[j + 1, j] == 0.0 or res <= tol:
                break
            V[j + 1] = w / H[j + 1, j]
        x = x + y @ V[:used]
        if res <= tol:
            return x, its


def strips(ntot, G):
    nb = ntot // G
    return [slice(k * nb, (k + 1) * nb) for k in range(G)]


def msm(A, b, G, rtol=1e-6, inner=None, max_outer=100000):
    """synchronous-multisplitting.c:170-206: block-Jacobi sweeps with capped inner GMRES; stops on the TRUE residual
    sqrt(sum_K ||rhs_K - A_KK x_K||^2) taken after the exchange and the rhs update."""
    inner = inner or {}
    sl = strips(b.size, G)
    AKK = [A[s, s].tocsr() for s in sl]
    x = np.zeros(b.size)
    norm0 = np.linalg.norm(b)
    for it in range(1, max_outer + 1):
        xn = x.copy()
        for s, a in zip(sl, AKK):
            rhs = b[s] - (A[s] @ x - a @ x[s])         # b_K - sum_J A_KJ x_J with the previous sweep's neighbours
            xn[s], _ = gmres_capped(a, rhs, x[s].copy(), **inner)
        x = xn
        if np.linalg.norm(b - A @ x) <= rtol * norm0:
            return x, it
    return x, max_outer


def smsm_global(A, b, G, s=5, rtol=1e-6, inner=None, max_outer=1000):
    """…-minimization-global.c:288-363: s block-Jacobi sweeps give S = [x^1 .. x^s]; alpha = argmin ||b - A S alpha||_2;
    x = S alpha; stop when that minimal residual <= rtol ||b||.  Returns (x, outer iterations, residual history)."""
    inner = inner or {}
    sl = strips(b.size, G)
    AKK = [A[q, q].tocsr() for q in sl]
    x = np.zeros(b.size)
    norm0 = np.linalg.norm(b)
    hist = []
    for it in range(1, max_outer + 1):
        S = np.zeros((b.size, s))
        for t in range(s):
            xn = x.copy()
            for q, a in zip(sl, AKK):
                rhs = b[q] - (A[q] @ x - a @ x[q])
                xn[q], _ = gmres_capped(a, rhs, x[q].copy(), **inner)
            x = xn
            S[:, t] = x
        R = A @ S
        alpha, *_ = np.linalg.lstsq(R, b, rcond=None)
        x = S @ alpha
        hist.append(np.linalg.norm(b - R @ alpha))
        if hist[-1] <= rtol * norm0:
            break
    return x, len(hist), np.array(hist)
