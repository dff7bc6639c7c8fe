"""Timing probe of the distributed inner solve (npb > 1) under torchrun: ms per Arnoldi step for a few sizes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from medane_tchakorom_ufc_thesis_repository_b200 import distributed as D  # noqa: E402
from medane_tchakorom_ufc_thesis_repository_b200 import solver as S  # noqa: E402

rank, world, local = D.env_rank()
torch.cuda.set_device(local)
npb = int(os.environ.get("PROBE_NPB", str(world)))
for (m, n, p, max_it, outer) in ((64, 64, 1, 20, 20), (2048, 2048, 1, 20, 10), (256, 256, 256, 20, 4)):
    eng = D.make_distributed_engine(m, n, p, s=0, max_restart=30, npb=npb)
    inner = S.ksp_opts(restart=30, max_it=max_it, rtol=1e-30, abstol=1e-300)
    eng.solve("SM", rtol=1e-300, inner=inner, max_outer=2)
    D.barrier()
    t0 = time.perf_counter()
    res = eng.solve("SM", rtol=1e-300, inner=inner, max_outer=outer)
    D.barrier()
    wall = time.perf_counter() - t0
    if rank == 0:
        steps = outer * max_it
        print(f"npb={npb} {m}x{n}x{p}: {res['elapsed_s'] * 1e3 / steps:.3f} ms per Arnoldi step (device), {wall * 1e3 / steps:.3f} wall, "
              f"{res['elapsed_s'] * 1e3 / outer:.2f} ms per inner solve, launches {res['kernel_launches']}", flush=True)
    eng.close()
    D.barrier()
