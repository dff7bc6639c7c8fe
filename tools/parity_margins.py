"""Prints, for every golden synchronous run (tests/golden/oracle_sync_runs.json), how far the device engine is from the CPU
oracle: after two outer iterations (x, history) and after the whole run (outer-iteration count, x, final residual).  The
bars of tests/test_gpu_parity.py::test_sync_driver_parity are set from this table (about ten times the measured deviation,
never below the north_star's 1e-8).  Run on a GPU: python tools/parity_margins.py > profiles/r02_parity_margins.txt"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from medane_tchakorom_ufc_thesis_repository_b200 import solver as S  # noqa: E402
from oracle import oracle as O  # noqa: E402

runs = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_sync_runs.json")))["runs"]
print(f"{'configuration':58s} {'dx(2 its)':>10s} {'dhist(2)':>10s} {'its gpu/orc':>12s} {'dx(run)':>10s} {'dresid(run)':>12s}")
for g in runs:
    args = dict(p=g.get("p", 1), nblocks=g["nblocks"], s=g["s"], inner=g["inner"])
    inner = S.ksp_opts(**g["inner"])
    grp = S.Group(g["m"], g["n"], g.get("p", 1), nblocks=g["nblocks"], s=g["s"], max_restart=g["inner"]["restart"])
    res = grp.solve(g["alg"], s=g["s"], rtol=1e-300, inner=inner, max_outer=2)
    ref = O.solve(g["alg"], g["m"], g["n"], rtol=1e-300, max_outer=2, **args)
    dx2 = np.linalg.norm(grp.solution() - ref["x"]) / np.linalg.norm(ref["x"])
    dh2 = np.max(np.abs(np.max([r["hist"] for r in res], axis=0) / ref["hist"] - 1.0))
    grp.close()
    grp = S.Group(g["m"], g["n"], g.get("p", 1), nblocks=g["nblocks"], s=g["s"], max_restart=g["inner"]["restart"])
    res = grp.solve(g["alg"], s=g["s"], rtol=g["rtol"], inner=inner, max_outer=3000)
    ref = O.solve(g["alg"], g["m"], g["n"], rtol=g["rtol"], max_outer=3000, **args)
    dxr = np.linalg.norm(grp.solution() - ref["x"]) / np.linalg.norm(ref["x"])
    dr = abs(res[0]["final_residual"] - ref["final_residual"]) / ref["final_residual"]
    grp.close()
    name = f"{g['alg']} {g['m']}x{g['n']}x{g.get('p', 1)} G={g['nblocks']} s={g['s']} max_it={g['inner']['max_it']} rtol={g['rtol']:g}"
    print(f"{name:58s} {dx2:10.2e} {dh2:10.2e} {res[0]['outer_its']:5d}/{ref['outer_its']:<5d} {dxr:10.2e} {dr:12.2e}", flush=True)
