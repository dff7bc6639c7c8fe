"""Residual history of SMSM-global at 1024x1024 (one block) next to the oracle run recorded in
tests/golden/smsm_global_1024_to_rtol.json: where the two trajectories decorrelate.  Needs a GPU."""
import sys, json, os
sys.path.insert(0, "/root/repo")
import numpy as np
from medane_tchakorom_ufc_thesis_repository_b200 import solver as S
gold = json.load(open("/root/repo/tests/golden/smsm_global_1024_to_rtol.json"))
e = S.Engine(1024, 1024, s=5, max_restart=30)
res = e.solve("SMSM_GLOBAL", s=5, rtol=1e-6, inner=S.ksp_opts(restart=30, max_it=20, rtol=1e-10, abstol=1e-100), max_outer=5000)
h, g = res["hist"], np.array(gold["hist"])
print("its", res["outer_its"], gold["outer_its"], "elapsed", res["elapsed_s"])
for i in list(range(0, min(len(h), len(g)), 8)):
    print(i, "%.6e %.6e rel diff %.2e" % (h[i], g[i], abs(h[i] - g[i]) / g[i]))
