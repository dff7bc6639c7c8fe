"""Small end-to-end run touching every kernel (for compute-sanitizer): sync + async drivers, TSQR and LSQR minimisers,
refinement/MGS variants, 2-D and 3-D, the persistent cycle kernel, stand-alone GMRES, raw op wrappers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from medane_tchakorom_ufc_thesis_repository_b200 import solver as S  # noqa: E402

inner = S.ksp_opts(restart=10, max_it=6, rtol=1e-10, abstol=1e-100)
for alg, s in (("SM", 0), ("SMSM_GLOBAL", 3), ("SMSM_SEMI_LOCAL", 3), ("SMSM_LOCAL", 3)):
    g = S.Group(24, 17, nblocks=3, s=s, max_restart=10)
    r = g.solve(alg, s=s, rtol=1e-4, inner=inner, max_outer=30)
    print(alg, r[0]["outer_its"], r[0]["final_residual"] / r[0]["norm0"])
    g.close()
# stencil-shaped strips (grid columns a multiple of 4): the coded-DIA stencil SpMV and, forced, the persistent cooperative
# restart-cycle kernel (csrc/cycle_coop.cuh) on three engines sharing the GPU, 2-D and 3-D
os.environ["MSPLIT_COOP"] = "1"
for shape, nblocks in (((24, 16, 1), 3), ((8, 4, 6), 2)):
    g = S.Group(*shape, nblocks=nblocks, s=3, max_restart=10)
    assert all(e.persistent_cycles() for e in g.engines)
    r = g.solve("SMSM_GLOBAL", s=3, rtol=1e-4, inner=inner, max_outer=30)
    print("persistent cycles", shape, r[0]["outer_its"], r[0]["final_residual"] / r[0]["norm0"])
    g.close()
del os.environ["MSPLIT_COOP"]
g = S.Group(6, 5, 8, nblocks=2, s=3, max_restart=10)
print("3d", g.solve("SMSM_GLOBAL", s=3, rtol=1e-4, inner=inner, max_outer=20)[0]["outer_its"])
g.close()
g = S.Group(16, 16, nblocks=2, s=3, max_restart=10)
print("lsqr", g.solve("SMSM_GLOBAL", s=3, rtol=1e-3, inner=inner, max_outer=10, outer_type="lsqr", outer_max_it=10)[0]["outer_its"])
g.close()
for alg, s in (("AM", 0), ("AMAM_GLOBAL", 3), ("AMAM_LOCAL", 3)):
    g = S.Group(16, 16, nblocks=2, s=s, max_restart=10)
    r = g.solve(alg, s=s, rtol=1e-3, inner=S.ksp_opts(restart=10, max_it=3, rtol=1e-10, abstol=1e-100), max_outer=400, periods=[1, 2])
    print(alg, [x["outer_its"] for x in r])
    g.close()
e = S.Engine(21, 19, max_restart=12, keep_csr=True)
for kw in (dict(cgs_refine=1), dict(cgs_refine=2), dict(mgs=1)):
    print("gmres", kw, e.gmres_solve(S.ksp_opts(restart=12, max_it=60, rtol=1e-6, abstol=1e-100, initial_rtol=1, **kw))["gmres_its"])
rng = np.random.default_rng(0)
V = rng.standard_normal((9, e.nb)); w = rng.standard_normal(e.nb)
print("ops", e.mdot(V, w)[:2], e.maxpy(V, -e.mdot(V, w), w)[1], e.spmv(S.MAT_DIAG, w)[:2])
print("csr", [len(a) for a in e.divideSubDomainIntoBlockMatrices(S.MAT_OFFDIAG)])
e.close()
print("SANITIZER TARGET DONE")
