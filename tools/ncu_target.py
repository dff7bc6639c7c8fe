"""Short target for ncu: one SMSM-global outer iteration (s=5, 100 Arnoldi steps + A*S + TSQR) at N x N on cuda:0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medane_tchakorom_ufc_thesis_repository_b200 import solver as S  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
if len(sys.argv) > 2 and sys.argv[2] == "3d":
    eng = S.Engine(N, N, N, s=5, max_restart=30)  # e.g. 256 -> 16.7 M rows, the per-GPU share of 512^3 on 8 GPUs
else:
    eng = S.Engine(N, N, s=5, max_restart=30)
res = eng.solve("SMSM_GLOBAL", s=5, rtol=1e-6, inner=S.ksp_opts(restart=30, max_it=20, rtol=1e-10, abstol=1e-100), max_outer=1)
print("outer_its", res["outer_its"], "launches", res["kernel_launches"], "elapsed_s", res["elapsed_s"])
eng.close()
