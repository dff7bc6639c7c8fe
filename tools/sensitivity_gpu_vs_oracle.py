"""relative ||x_gpu - x_oracle|| after 1/2/3 outer iterations for each minimisation variant,
block count and inner max_it (32x32).  Needs a GPU."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from medane_tchakorom_ufc_thesis_repository_b200 import solver as S
from oracle import oracle as O
for alg in ("SMSM_GLOBAL", "SMSM_SEMI_LOCAL", "SMSM_LOCAL"):
  for G in (1, 2, 4):
    for mi in (5, 20):
        inner = dict(restart=30, max_it=mi, rtol=1e-10, abstol=1e-100)
        out = []
        for mo in (1, 2, 3):
            grp = S.Group(32, 32, nblocks=G, s=5, max_restart=30)
            res = grp.solve(alg, s=5, rtol=1e-12, inner=S.ksp_opts(**inner), max_outer=mo)
            ref = O.solve(alg, 32, 32, nblocks=G, s=5, rtol=1e-12, inner=inner, max_outer=mo)
            x = grp.solution()
            out.append("%.1e" % (np.linalg.norm(x - ref["x"]) / np.linalg.norm(ref["x"])))
            grp.close()
        print(alg, "G", G, "max_it", mi, out, flush=True)
