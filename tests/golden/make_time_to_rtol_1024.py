"""Oracle run of the headline algorithm to convergence: SMSM-global s=5, 2-D 1024x1024, ONE block, inner GMRES(30)
max_it 20 rtol 1e-10, rtol 1e-6 (BASELINE configs[2] on a grid where time-to-rtol is reachable)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O
t = time.time()
r = O.solve("SMSM_GLOBAL", 1024, 1024, nblocks=1, s=5, rtol=1e-6, inner=dict(restart=30, max_it=20, rtol=1e-10, abstol=1e-100),
            max_outer=2000, nthreads=8, want_x=False)
json.dump({"generator": "tests/golden/make_time_to_rtol_1024.py", "outer_its": r["outer_its"], "norm0": r["norm0"],
           "final_residual": r["final_residual"], "error": r["error"], "hist": list(r["hist"]), "oracle_seconds_8threads": time.time() - t},
          open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "smsm_global_1024_to_rtol.json"), "w"))
print(r["outer_its"], r["final_residual"] / r["norm0"], time.time() - t)
