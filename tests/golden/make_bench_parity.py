"""Generates tests/golden/bench_parity_G{2,4,8}.npz: what bench.py's `parity_check` (run under torchrun before the timed
region) and tools/mgpu_check.py compare the one-process-per-GPU path against.  Fixtures OF THE ORACLE
(the reference cannot run here): SMSM-global, 64x64, s = 5, inner GMRES(30) capped at 5 — the well-separated regime where
GPU and oracle agree to rounding — x and the residual history after 3 outer iterations, and the outer-iteration count of
the run to rtol 1e-6."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O  # noqa: E402

inner = dict(restart=30, max_it=5, rtol=1e-10, abstol=1e-100)
here = os.path.dirname(os.path.abspath(__file__))
for G in (1, 2, 4, 8):  # G = 1: what a run with ONE Jacobi block spread over all GPUs (-npb N) must reproduce
    r3 = O.solve("SMSM_GLOBAL", 64, 64, nblocks=G, s=5, rtol=1e-300, inner=inner, max_outer=3)
    rf = O.solve("SMSM_GLOBAL", 64, 64, nblocks=G, s=5, rtol=1e-6, inner=inner, max_outer=5000, want_x=False)
    r5 = O.solve("SMSM_GLOBAL", 64, 64, nblocks=G, s=5, rtol=1e-5, inner=inner, max_outer=5000, want_x=False)
    np.savez(os.path.join(here, f"bench_parity_G{G}.npz"), x3=r3["x"], hist3=r3["hist"], norm0=r3["norm0"],
             outer_its_to_1e6=rf["outer_its"], outer_its_to_1e5=r5["outer_its"], final_rel=rf["final_residual"] / rf["norm0"])
    print(G, r3["hist"], rf["outer_its"], r5["outer_its"])
