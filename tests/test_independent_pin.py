"""Pins the oracle's SOLVER arithmetic (GMRES-capped block-Jacobi, exact least-squares minimisation, LSQR) against an
independent numpy/scipy implementation (tests/independent_reference.py) and against scipy's own LSQR.

The reference's Unity tests pin assembly and the residual norm only; PETSc is not on disk.  Two
implementations that share no code and no algorithmic shortcut (Givens recurrence vs. lstsq on the Hessenberg matrix,
Householder QR on the correction basis vs. lstsq on the raw iterates) agreeing on iteration counts and residual histories
is the strongest pin available here.  CPU only.
"""
import numpy as np
import pytest

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import independent_reference as I  # noqa: E402
from oracle import oracle as O  # noqa: E402

INNER20 = dict(restart=30, max_it=20, rtol=1e-10, abstol=1e-100)
INNER50 = dict(restart=30, max_it=50, rtol=1e-10, abstol=1e-100)
INNER5 = dict(restart=30, max_it=5, rtol=1e-10, abstol=1e-100)


def _ind(inner):
    return dict(restart=inner["restart"], max_it=inner["max_it"], rtol=inner["rtol"], abstol=inner["abstol"])


def test_matrix_equals_oracle_assembly():
    for (m, n) in ((6, 5), (16, 16)):
        rp, ci, va = O.poisson2d(m, n)
        A = I.poisson2d(m, n)
        A.sort_indices()
        assert np.array_equal(A.indptr, rp) and np.array_equal(A.indices, ci) and np.array_equal(A.data, va)
    rp, ci, va = O.poisson3d(4, 3, 5)
    A = I.poisson3d(4, 3, 5)
    A.sort_indices()
    assert np.array_equal(A.indptr, rp) and np.array_equal(A.indices, ci) and np.array_equal(A.data, va)


def test_capped_gmres_matches_oracle_iterate():
    """One inner solve: same iteration count, same iterate (1e-10), for a cut-short cycle, a restart and a converged run."""
    A = I.poisson2d(24, 20)
    b = A @ np.ones(A.shape[0])
    rp, ci, va = O.poisson2d(24, 20)
    rng = np.random.default_rng(7)
    x0 = rng.standard_normal(A.shape[0])
    for kw in (dict(restart=30, max_it=7, rtol=1e-10), dict(restart=5, max_it=23, rtol=1e-10), dict(restart=30, max_it=10000, rtol=1e-6)):
        xi, its_i = I.gmres_capped(A, b, x0.copy(), abstol=1e-100, **kw)
        xo, its_o, reason, rn = O.gmres(rp, ci, va, b, x0=x0, abstol=1e-100, initial_rtol=1, guess_nonzero=1, **kw)
        assert its_i == its_o, (kw, its_i, its_o)
        assert np.linalg.norm(xi - xo) <= 1e-9 * np.linalg.norm(xo)


@pytest.mark.parametrize("refine", [1, 2])
def test_gmres_with_cgs_refinement_matches_oracle(refine):
    """-ksp_gmres_cgs_refinement_type refine_ifneeded / refine_always: same iteration count and iterate as the oracle."""
    A = I.poisson2d(28, 24)
    b = A @ np.ones(A.shape[0])
    rp, ci, va = O.poisson2d(28, 24)
    x0 = np.random.default_rng(11).standard_normal(A.shape[0])
    for kw in (dict(restart=30, max_it=45, rtol=1e-12), dict(restart=12, max_it=10000, rtol=1e-8)):
        xi, its_i = I.gmres_capped(A, b, x0.copy(), abstol=1e-100, refine=refine, **kw)
        xo, its_o, reason, rn = O.gmres(rp, ci, va, b, x0=x0, abstol=1e-100, initial_rtol=1, guess_nonzero=1, cgs_refine=refine, **kw)
        assert its_i == its_o, (kw, its_i, its_o)
        assert np.linalg.norm(xi - xo) <= 1e-9 * np.linalg.norm(xo)


def test_smsm_global_one_block_256():
    """VERDICT r01 point 1: 256^2, one block, s = 5, inner GMRES(30) capped at 20 -> 12 outer iterations, 6.07e-7."""
    n = 256
    A = I.poisson2d(n, n)
    b = A @ np.ones(n * n)
    x, its, hist = I.smsm_global(A, b, 1, s=5, rtol=1e-6, inner=_ind(INNER20))
    ref = O.solve("SMSM_GLOBAL", n, n, nblocks=1, s=5, rtol=1e-6, inner=INNER20)
    assert its == ref["outer_its"] == 12
    assert abs(hist[-1] / np.linalg.norm(b) - 6.07e-7) < 0.01e-7
    assert np.allclose(hist, ref["hist"], rtol=1e-6)
    assert np.linalg.norm(x - ref["x"]) <= 1e-8 * np.linalg.norm(x)


def test_msm_two_blocks_sweep_count():
    """VERDICT r01 point 2 (at a size the CPU suite can afford): MSM, 2 blocks, inner max_it 50: sweep count +-1."""
    m = n = 96
    A = I.poisson2d(m, n)
    b = A @ np.ones(m * n)
    x, its = I.msm(A, b, 2, rtol=1e-6, inner=_ind(INNER50))
    ref = O.solve("SM", m, n, nblocks=2, s=0, rtol=1e-6, inner=INNER50)
    assert abs(its - ref["outer_its"]) <= 1, (its, ref["outer_its"])
    assert np.linalg.norm(x - ref["x"]) <= 1e-5 * np.linalg.norm(x)  # both stopped at rel. residual 1e-6


def test_smsm_global_two_blocks_no_convergence_in_40():
    """VERDICT r01 point 3: 256^2, 2 blocks, inner max_it 20.  The first outer iterations agree value for value; after 40
    outer iterations NEITHER implementation has reached 1e-6.  Where each one flattens depends on how its least-squares
    solver treats the nearly dependent basis columns (measured: Householder QR on the correction basis 1.6e-4, PETSc-style
    LSQR 9e-5, numpy.linalg.lstsq's SVD with its default cut-off 9e-6) — the regime DESIGN.md §5 calls chaotic, which
    is why whole-run parity is asserted on iteration counts of CONVERGING configurations only."""
    n = 256
    A = I.poisson2d(n, n)
    b = A @ np.ones(n * n)
    nrm = np.linalg.norm(b)
    x, its, hist = I.smsm_global(A, b, 2, s=5, rtol=1e-6, inner=_ind(INNER20), max_outer=40)
    ref = O.solve("SMSM_GLOBAL", n, n, nblocks=2, s=5, rtol=1e-6, inner=INNER20, max_outer=40)
    assert its == ref["outer_its"] == 40                     # neither converges
    assert np.allclose(hist[:3], ref["hist"][:3], rtol=1e-7)  # before rounding differences are amplified
    assert hist[-1] / nrm > 2e-6 and ref["hist"][-1] / nrm > 2e-6
    assert 0.5e-4 < ref["hist"][-1] / nrm < 4e-4            # the level VERDICT r01 measured independently (1.7e-4 .. 1.8e-4)


@pytest.mark.parametrize("G,s,inner,k_hist", [(4, 5, INNER5, 3), (2, 4, INNER20, 1), (8, 10, INNER20, 3)])
def test_smsm_global_3d_iteration_count(G, s, inner, k_hist):
    """3-D 7-point strips (the north_star problem at a small size): outer-iteration count +-1 and the residual history
    (one outer iteration only where two big blocks with near-exact inner solves make the iterates collinear: the
    deviation there grows 1e-12 -> 2e-4 -> 7e-2 over the first three, the amplification DESIGN.md §5 describes)."""
    N = 16
    A = I.poisson3d(N, N, N)
    b = A @ np.ones(N ** 3)
    x, its, hist = I.smsm_global(A, b, G, s=s, rtol=1e-6, inner=_ind(inner))
    ref = O.solve("SMSM_GLOBAL", N, N, N, nblocks=G, s=s, rtol=1e-6, inner=inner)
    assert abs(its - ref["outer_its"]) <= 1, (its, ref["outer_its"])
    k = min(k_hist, its, ref["outer_its"])
    assert np.allclose(hist[:k], ref["hist"][:k], rtol=1e-6)


def test_orc_lsqr_matches_scipy_lsqr():
    """KSPSolve_LSQR restatement vs scipy.sparse.linalg.lsqr (both Paige-Saunders): the iterate after k iterations and the
    residual estimate phibar."""
    from scipy.sparse.linalg import lsqr
    rng = np.random.default_rng(3)
    R = rng.standard_normal((400, 6)) @ np.diag([1, 0.5, 0.1, 1e-2, 1e-3, 1e-4])
    b = rng.standard_normal(400)
    for k in (3, 6, 12):
        xs = lsqr(R, b, atol=0.0, btol=0.0, conlim=0.0, iter_lim=k)
        alpha, its, reason, rnorm = O.lsqr(R, b, max_it=k, rtol=1e-300, abstol=1e-300, default_test=True)
        assert its == k
        assert np.linalg.norm(alpha - xs[0]) <= 1e-7 * np.linalg.norm(xs[0])  # measured 3e-14 / 6e-9 / 4e-14 (k = 6 resolves the 1e-4 column)
        assert abs(rnorm - xs[3]) <= 1e-10 * xs[3]   # r1norm = phibar
