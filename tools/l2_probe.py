"""One-GPU proxy for the per-GPU share of the 8-GPU runs: seconds per SMSM-global outer iteration on an M x N grid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medane_tchakorom_ufc_thesis_repository_b200 import solver as S  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "8192x1024"
dims = [int(t) for t in shape.split("x")]
outer = int(sys.argv[2]) if len(sys.argv) > 2 else 4
eng = S.Engine(*dims, s=5, max_restart=30)
inner = S.ksp_opts(restart=30, max_it=20, rtol=1e-10, abstol=1e-100)
eng.solve("SMSM_GLOBAL", s=5, rtol=1e-300, inner=inner, max_outer=2)
res = eng.solve("SMSM_GLOBAL", s=5, rtol=1e-300, inner=inner, max_outer=outer)
print(shape, "env", {k: v for k, v in os.environ.items() if k.startswith("MSPLIT")}, "ms per outer iteration", 1e3 * res["elapsed_s"] / res["outer_its"],
      "hist", res["hist"][-1] if res.get("hist") is not None and len(res["hist"]) else None, flush=True)
eng.close()
