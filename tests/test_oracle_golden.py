"""Pins the CPU oracle against every known answer the reference's own unit tests hold for this path
(src/tests/utils_test.c, run there as `mpirun -n 4 ./bin/utils_test`: 2 blocks x 2 ranks)."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def dense_rows(rp, ci, va, ncols):
    n = len(rp) - 1
    D = np.zeros((n, ncols))
    for r in range(n):
        for k in range(rp[r], rp[r + 1]):
            D[r, ci[k]] = va[k]
    return D


def test_computeDimensionRelatedVariables(oracle):
    # utils_test.c:38-64 — 2x2 mesh, 4 ranks, npb = 2
    exp = {0: (0, 0), 1: (1, 0), 2: (0, 1), 3: (1, 1)}
    for rank in range(4):
        d = oracle.dimension_related(4, 2, rank, 2, 2)
        assert d["njacobi_blocks"] == 2 and d["n_mesh_points"] == 4 and d["jacobi_block_size"] == 2
        assert (d["proc_local_rank"], d["rank_jacobi_block"]) == exp[rank]


def test_poisson2DMatrix(oracle):
    # utils_test.c:172-221
    D0 = dense_rows(*oracle.poisson2d(2, 2, 0, 2), 4)
    D1 = dense_rows(*oracle.poisson2d(2, 2, 1, 2), 4)
    assert np.array_equal(D0, [[4, -1, -1, 0], [-1, 4, 0, -1]])
    assert np.array_equal(D1, [[-1, 0, 4, -1], [0, -1, -1, 4]])


def test_poisson3DMatrix(oracle):
    # utils_test.c:66-170 — 2x2x2 mesh: diag 6, -1 at +-1, +-2, +-4 where in range
    D = np.vstack([dense_rows(*oracle.poisson3d(2, 2, 2, k, 2), 8) for k in range(2)])
    exp = np.zeros((8, 8))
    for r in range(8):
        i, j, k = r % 2, (r // 2) % 2, r // 4
        exp[r, r] = 6
        if i > 0: exp[r, r - 1] = -1
        if i < 1: exp[r, r + 1] = -1
        if j > 0: exp[r, r - 2] = -1
        if j < 1: exp[r, r + 2] = -1
        if k > 0: exp[r, r - 4] = -1
        if k < 1: exp[r, r + 4] = -1
    assert np.array_equal(D, exp)
    with open(os.path.join(GOLD, "utils_test_poisson3d_2x2x2.json")) as f:
        gold = json.load(f)
    assert np.array_equal(D, np.array(gold["rows"]))


def test_computeFinalResidualNorm(oracle):
    # utils_test.c:225-228 with inputs :285-316 => 2.54567588 (TEST_ASSERT_EQUAL_FLOAT)
    x0 = np.array([0.1234, 0.5678, 0.9101, 0.1121]); b0 = np.array([0.3141, 0.5926])
    x1 = np.array([0.8765, 0.4321, 0.5432, 0.6789]); b1 = np.array([0.2468, 0.1357])
    n0 = oracle.block_residual_norm(*oracle.poisson2d(2, 2, 0, 2), b0, x0)
    n1 = oracle.block_residual_norm(*oracle.poisson2d(2, 2, 1, 2), b1, x1)
    got = np.sqrt(n0 * n0 + n1 * n1)
    assert abs(np.float32(got) - np.float32(2.54567588)) <= 1e-5 * 2.54567588
    assert abs(got - 2.5456758807829405) < 1e-14


def test_csr_sorted_and_counts(oracle):
    # nnz_total = 5mn - 2m - 2n (2-D), 7N^3 - 6N^2 (3-D)  
    for (m, n, G) in [(8, 8, 2), (6, 10, 3), (16, 4, 4)]:
        tot = 0
        for k in range(G):
            rp, ci, va = oracle.poisson2d(m, n, k, G)
            tot += rp[-1]
            for r in range(len(rp) - 1):
                row = ci[rp[r]:rp[r + 1]]
                assert np.all(np.diff(row) > 0)
        assert tot == 5 * m * n - 2 * m - 2 * n
    N = 6
    tot = sum(oracle.poisson3d(N, N, N, k, 2)[0][-1] for k in range(2))
    assert tot == 7 * N ** 3 - 6 * N ** 2


def test_split_matches_scipy(oracle):
    import scipy.sparse as sp
    m, n, G, K = 8, 6, 2, 1
    rp, ci, va = oracle.poisson2d(m, n, K, G)
    nb = m * n // G
    A = sp.csr_matrix((va, ci, rp), shape=(nb, m * n))
    drp, dci, dva = oracle.submatrix(rp, ci, va, K * nb, (K + 1) * nb)
    D = sp.csr_matrix((dva, dci, drp), shape=(nb, nb))
    assert (abs(D - A[:, K * nb:(K + 1) * nb]).sum()) == 0


def test_gmres_matches_scipy_solution(oracle):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    n = 24
    rp, ci, va = oracle.poisson2d_complete(n, n)
    A = sp.csr_matrix((va, ci, rp), shape=(n * n, n * n))
    b = A @ np.ones(n * n)
    x, its, reason, rnorm = oracle.gmres(rp, ci, va, b, restart=30, rtol=1e-10, max_it=2000, initial_rtol=1)
    assert reason == 2
    assert np.linalg.norm(b - A @ x) <= 1.5e-10 * np.linalg.norm(b)
    assert abs(rnorm - np.linalg.norm(b - A @ x)) <= 1e-6 * rnorm + 1e-14
    xs = spla.spsolve(A.tocsc(), b)
    assert np.linalg.norm(x - xs) < 1e-7


def test_gmres_max_it_cuts_cycle_and_min_it_rule(oracle):
    n = 16
    rp, ci, va = oracle.poisson2d_complete(n, n)
    b = oracle.spmv(rp, ci, va, np.ones(n * n))
    # max_it = 7 < restart: exactly 7 iterations, DIVERGED_ITS (-3), iterate still updated
    x, its, reason, _ = oracle.gmres(rp, ci, va, b, restart=30, rtol=1e-30, max_it=7)
    assert its == 7 and reason == -3 and np.linalg.norm(x) > 0
    # starting from the exact solution: residual 0 on entry => CONVERGED_ATOL, 0 iterations
    x, its, reason, _ = oracle.gmres(rp, ci, va, b, x0=np.ones(n * n), restart=30, guess_nonzero=1)
    assert its == 0 and reason == 3


def test_lsqr_and_qr_agree(oracle):
    rng = np.random.default_rng(7)
    R = rng.standard_normal((200, 5))
    b = rng.standard_normal(200)
    a_ref, *_ = np.linalg.lstsq(R, b, rcond=None)
    a_qr, rn_qr = oracle.lstsq_qr(R, b)
    a_ls, its, reason, rn_ls = oracle.lsqr(R, b, max_it=70, rtol=1e-15)
    assert np.allclose(a_qr, a_ref, rtol=1e-10, atol=1e-12)
    assert np.allclose(a_ls, a_ref, rtol=1e-8, atol=1e-10)
    assert abs(rn_qr - np.linalg.norm(b - R @ a_ref)) < 1e-10
    assert abs(rn_ls - rn_qr) < 1e-8
    assert its == 70  # inconsistent system + rtol 1e-15: LSQR always runs max_it


def test_sync_drivers_converge(oracle):
    inner = dict(restart=30, max_it=20, rtol=1e-10, abstol=1e-100)
    r = oracle.solve("SM", 32, 32, nblocks=2, rtol=1e-6, inner=dict(restart=30, max_it=50, rtol=1e-10, abstol=1e-100))
    assert r["rc"] == 0 and r["final_residual"] <= 1e-6 * r["norm0"] and r["outer_its"] == 108
    r = oracle.solve("SMSM_GLOBAL", 32, 32, nblocks=2, s=5, rtol=1e-6, inner=inner)
    assert r["rc"] == 0 and r["outer_its"] == 6 and r["final_residual"] <= 1e-6 * r["norm0"]
    rl = oracle.solve("SMSM_GLOBAL", 32, 32, nblocks=2, s=5, rtol=1e-6, inner=inner, outer_type="lsqr", outer_max_it=70)
    assert abs(rl["outer_its"] - r["outer_its"]) <= 1
    for alg in ("SMSM_SEMI_LOCAL", "SMSM_LOCAL"):
        r = oracle.solve(alg, 32, 32, nblocks=2, s=5, rtol=1e-6, inner=inner)
        assert r["rc"] == 0 and r["outer_its"] > 0 and r["last_norm"] <= 1e-6 / np.sqrt(2) * r["norm0"]
    with open(os.path.join(GOLD, "oracle_sync_runs.json")) as f:
        gold = json.load(f)
    for g in gold["runs"]:
        r = oracle.solve(g["alg"], g["m"], g["n"], p=g.get("p", 1), nblocks=g["nblocks"], s=g["s"], rtol=g["rtol"],
                         inner=g["inner"])
        assert r["outer_its"] == g["outer_its"], g
        assert abs(r["final_residual"] - g["final_residual"]) <= 1e-6 * g["final_residual"]


def test_conv_detection_two_blocks(oracle):
    # both roots under the threshold with fresh data every step => elected leader = max rank, positive verdict
    cd = oracle.ConvDetect(2)
    it = [0, 0]
    finished = False
    for step in range(40):
        for k in (0, 1):
            other = 1 - k
            cd.data_arrival(k, other, cd.phase_tag(other), it[other])
            cd.step(k, True)
            it[k] += 1
        if cd.state(0) == cd.FINISHED and cd.state(1) == cd.FINISHED:
            finished = True
            break
    assert finished
    # never under the threshold => stays NORMAL
    cd = oracle.ConvDetect(2)
    for step in range(20):
        for k in (0, 1):
            cd.data_arrival(k, 1 - k, 0, step)
            cd.step(k, False)
    assert cd.state(0) == cd.NORMAL and cd.state(1) == cd.NORMAL


def test_async_drivers_reach_residual(oracle):
    # inexact inner solves (max_it 3): the local residual then tracks the global one and the detection
    # protocol (pseudo-periods, election, verification, verdict) stops near the requested tolerance
    inner = dict(restart=30, max_it=3, rtol=1e-10, abstol=1e-100)
    r = oracle.solve("AM", 24, 24, nblocks=2, rtol=1e-5, inner=inner, periods=[1, 2], max_outer=4000)
    assert r["rc"] == 0 and r["final_residual"] <= 2e-5 * r["norm0"]
    assert r["outer_its_block"][0] == 2 * r["outer_its_block"][1]  # block 1 runs every second tick
    for alg in ("AMAM_GLOBAL", "AMAM_SEMI_LOCAL", "AMAM_LOCAL"):
        r = oracle.solve(alg, 24, 24, nblocks=2, s=4, rtol=1e-5, inner=inner, periods=[1, 2], max_outer=4000)
        assert r["rc"] == 0 and r["final_residual"] <= 1e-4 * r["norm0"], alg


def test_async_four_blocks_chain_terminates(oracle):
    """Generalisation of the 2-block detector to a chain of block roots: uneven speeds, 4 blocks,
    2-D and 3-D; every block reaches FINISHED through partial-CV / verification / response / verdict messages."""
    inner = dict(restart=30, max_it=3, rtol=1e-10, abstol=1e-100)
    r = oracle.solve("AM", 32, 16, nblocks=4, rtol=1e-4, inner=inner, periods=[1, 2, 1, 3], max_outer=6000)
    assert r["rc"] == 0 and all(i > 0 for i in r["outer_its_block"]) and r["final_residual"] <= 1e-3 * r["norm0"]
    r = oracle.solve("AMAM_GLOBAL", 8, 8, p=8, nblocks=4, s=3, rtol=1e-4, inner=inner, periods=[2, 1, 1, 1], max_outer=6000)
    assert r["rc"] == 0 and r["final_residual"] <= 1e-3 * r["norm0"]


def test_sync_run_fixtures_are_what_the_oracle_produces(oracle):
    """tests/golden/oracle_sync_runs.json (made by make_oracle_sync_runs.py) is what the GPU parity tests compare against:
    re-run a subset here so that a change of the oracle cannot silently move the bar."""
    with open(os.path.join(GOLD, "oracle_sync_runs.json")) as f:
        runs = json.load(f)["runs"]
    checked = 0
    for g in runs:
        if g["outer_its"] > 60:      # keep the CPU suite short: the long SM sweeps are re-run by the GPU tests anyway
            continue
        r = oracle.solve(g["alg"], g["m"], g["n"], p=g.get("p", 1), nblocks=g["nblocks"], s=g["s"], rtol=g["rtol"], inner=g["inner"],
                         max_outer=3000)
        assert r["outer_its"] == g["outer_its"], (g["alg"], g["m"], g["n"])
        assert r["inner_its_total"] == g["inner_its_total"]
        assert abs(r["norm0"] - g["norm0"]) <= 1e-14 * g["norm0"]
        assert abs(r["final_residual"] - g["final_residual"]) <= 1e-9 * g["final_residual"]
        checked += 1
    assert checked >= 10


def test_assembly_properties_random_shapes(oracle):
    """Property test over random grid shapes and block counts (hypothesis): the strips stacked over the blocks are the
    symmetric stencil matrix with the closed-form row sums, and — the precondition of the engine's coded-DIA view —
    every diagonal of a strip holds ONE value wherever it is present."""
    import scipy.sparse as sp
    from hypothesis import given, settings, strategies as st

    def check(strips, ntot, expect_rowsum, diag_value):
        A = sp.vstack([sp.csr_matrix((va, ci, rp), shape=(len(rp) - 1, ntot)) for rp, ci, va in strips]).tocsr()
        assert A.shape == (ntot, ntot)
        assert abs(A - A.T).sum() == 0
        assert np.array_equal(np.asarray(A.sum(axis=1)).ravel(), expect_rowsum)
        assert np.all(A.diagonal() == diag_value)
        coo = A.tocoo()
        for off in np.unique(coo.col - coo.row):
            vals = coo.data[(coo.col - coo.row) == off]
            assert np.all(vals == vals[0])       # one constant per diagonal
        for rp, ci, va in strips:
            for r in range(len(rp) - 1):
                assert np.all(np.diff(ci[rp[r]:rp[r + 1]]) > 0)   # sorted columns, as MatAssembly leaves them

    @settings(max_examples=25, deadline=None)
    @given(st.integers(1, 4), st.integers(1, 6), st.integers(1, 12))
    def two_d(G, lines_per_block, n):
        m = G * lines_per_block
        idx = np.arange(m * n)
        i, j = idx // n, idx % n
        # one unit per missing neighbour (a 1-wide grid misses both column neighbours)
        rowsum = (i == 0).astype(float) + (i == m - 1) + (j == 0) + (j == n - 1)
        check([oracle.poisson2d(m, n, k, G) for k in range(G)], m * n, rowsum, 4.0)

    @settings(max_examples=15, deadline=None)
    @given(st.integers(1, 3), st.integers(1, 3), st.integers(1, 6), st.integers(1, 6))
    def three_d(G, planes_per_block, nx, ny):
        nz = G * planes_per_block
        idx = np.arange(nx * ny * nz)
        i, j, k = idx % nx, (idx // nx) % ny, idx // (nx * ny)
        rowsum = (i == 0).astype(float) + (i == nx - 1) + (j == 0) + (j == ny - 1) + (k == 0) + (k == nz - 1)
        check([oracle.poisson3d(nx, ny, nz, b, G) for b in range(G)], nx * ny * nz, rowsum, 6.0)

    two_d()
    three_d()
