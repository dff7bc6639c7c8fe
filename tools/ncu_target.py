"""Short target for ncu: one SMSM-global outer iteration (s basis vectors: s x 20 Arnoldi steps + A*S + minimisation) on cuda:0.
usage: python tools/ncu_target.py N [3d] [s]     (N x N, or N^3 with `3d`; s defaults to 5)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medane_tchakorom_ufc_thesis_repository_b200 import solver as S  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
three_d = len(sys.argv) > 2 and sys.argv[2] == "3d"
s = int(sys.argv[3]) if len(sys.argv) > 3 else 5
eng = S.Engine(N, N, N, s=s, max_restart=30) if three_d else S.Engine(N, N, s=s, max_restart=30)  # 256 3d -> 16.7 M rows, the per-GPU share of 512^3 on 8 GPUs
res = eng.solve("SMSM_GLOBAL", s=s, rtol=1e-6, inner=S.ksp_opts(restart=30, max_it=20, rtol=1e-10, abstol=1e-100), max_outer=1)
print("outer_its", res["outer_its"], "launches", res["kernel_launches"], "elapsed_s", res["elapsed_s"])
eng.close()
