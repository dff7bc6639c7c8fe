"""How sensitive is the outer-iteration count of SMSM-global to a rounding-level perturbation?  Oracle only (CPU):
b is multiplied by (1 + eps sin(i)), eps = 0, +-1e-15, 3e-15 (ORC_PERTURB_B).  Output: DESIGN.md §5."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1
code = f'''
import sys; sys.path.insert(0, {ROOT!r})
from oracle import oracle as O
r = O.solve("SMSM_GLOBAL", {N}, {N}, nblocks={G}, s=5, rtol=1e-6, inner=dict(restart=30, max_it=20, rtol=1e-10, abstol=1e-100), max_outer=3000, nthreads=8, want_x=False)
print(r["outer_its"], r["final_residual"] / r["norm0"])
'''
for eps in ("0", "1e-15", "-1e-15", "3e-15", "1e-14"):
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, ORC_PERTURB_B=eps), capture_output=True, text=True)
    print(f"N={N} G={G} eps={eps}: outer_its, rel residual = {out.stdout.strip()}", flush=True)
