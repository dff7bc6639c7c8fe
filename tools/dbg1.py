import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from medane_tchakorom_ufc_thesis_repository_b200 import solver as S
from oracle import oracle as O
import sys as _s
inner = dict(restart=30, max_it=20, rtol=1e-10, abstol=1e-100, cgs_refine=int(_s.argv[1]) if len(_s.argv)>1 else 0)
for mo in (1,2,3,6):
    grp = S.Group(32, 32, nblocks=2, s=5, max_restart=30)
    res = grp.solve("SMSM_GLOBAL", s=5, rtol=1e-6, inner=S.ksp_opts(**inner), max_outer=mo)
    ref = O.solve("SMSM_GLOBAL", 32, 32, nblocks=2, s=5, rtol=1e-6, inner=inner, max_outer=mo)
    x = grp.solution()
    print(mo, res[0]["outer_its"], ref["outer_its"], np.linalg.norm(x-ref["x"])/np.linalg.norm(ref["x"]), res[0]["hist"], ref["hist"], res[0]["final_residual"], ref["final_residual"])
    grp.close()
# MSM for comparison
grp = S.Group(32, 32, nblocks=2, max_restart=30)
res = grp.solve("SM", rtol=1e-6, inner=S.ksp_opts(**inner), max_outer=3)
ref = O.solve("SM", 32, 32, nblocks=2, rtol=1e-6, inner=inner, max_outer=3)
print("SM", np.linalg.norm(grp.solution()-ref["x"])/np.linalg.norm(ref["x"]), res[0]["hist"], ref["hist"])
