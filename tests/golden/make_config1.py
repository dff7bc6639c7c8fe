"""BASELINE config 1 on the CPU oracle: MSM 512x512, 2 blocks, rtol 1e-6, inner GMRES(30) max_it 50 rtol 1e-10
(running_bulk_test_local:96-101).  Writes config1_msm_512.json + a 4096-point sample of the solution."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O
t = time.time()
r = O.solve("SM", 512, 512, nblocks=2, rtol=1e-6, inner=dict(restart=30, max_it=50, rtol=1e-10, abstol=1e-100), max_outer=20000, nthreads=8)
el = time.time() - t
here = os.path.dirname(os.path.abspath(__file__))
idx = np.linspace(0, 512 * 512 - 1, 4096).astype(np.int64)
np.save(os.path.join(here, "config1_msm_512_x_sample.npy"), r["x"][idx])
json.dump({"generator": "tests/golden/make_config1.py", "outer_its": r["outer_its"], "inner_its_total": r["inner_its_total"],
           "norm0": r["norm0"], "final_residual": r["final_residual"], "error": r["error"], "oracle_seconds_8threads": el,
           "sample_idx": idx.tolist()}, open(os.path.join(here, "config1_msm_512.json"), "w"))
print(r["outer_its"], r["final_residual"] / r["norm0"], el)
