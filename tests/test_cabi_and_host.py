"""CPU-side checks: the C-ABI library loads and exports every symbol include/msplit.h declares (no compute
calls without a GPU), the host logic (partition, option parsing) and the multi-rank plumbing on gloo."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def test_library_loads_and_exports_header_symbols():
    from medane_tchakorom_ufc_thesis_repository_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "msplit.h")).read()
    declared = set(re.findall(r"\b(msp_[A-Za-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for sym in declared:
        assert hasattr(L, sym), sym
    assert L.msp_version() == 200


def test_no_cpu_fallback_message(tmp_path, monkeypatch):
    from medane_tchakorom_ufc_thesis_repository_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(_lib.MsplitError, match="no CPU fallback"):
        _lib.lib()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "medane_tchakorom_ufc_thesis_repository_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                for bad in ('#include "msplit_oracle', "#include <msplit_oracle", "libmsplit_oracle", "from oracle", "import oracle"):
                    assert bad not in src, (f, bad)


def test_partition_helpers():
    from medane_tchakorom_ufc_thesis_repository_b200 import distributed as D
    assert D.strip_partition(512, 2) == [(0, 256), (256, 512)]
    assert D.strip_partition(512, 8)[-1] == (448, 512)
    with pytest.raises(ValueError):
        D.strip_partition(10, 3)
    assert D.neighbours(0, 4) == [None, 1] and D.neighbours(3, 4) == [2, None] and D.neighbours(0, 1) == [None, None]
    # computeDimensionRelatedVariables utils.c:657-659 (utils_test.c:38-64)
    assert [D.block_of_rank(r, 4, 2) for r in range(4)] == [(0, 0), (0, 1), (1, 0), (1, 1)]


def test_dimension_related_cabi():
    from medane_tchakorom_ufc_thesis_repository_b200 import solver as S
    d = S.computeDimensionRelatedVariables(4, 2, 3, 2, 2)
    assert d == {"njacobi_blocks": 2, "rank_jacobi_block": 1, "proc_local_rank": 1, "n_mesh_points": 4, "jacobi_block_size": 2}


_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["MSP_ROOT"])
import torch.distributed as dist
from medane_tchakorom_ufc_thesis_repository_b200 import distributed as D
rank, world, local = D.env_rank()
D.init_process_group("gloo")
calls = []
D.connect_blocks(rank, world, lambda: b"window-of-%d" % rank, lambda side, h: calls.append((side, h)),
                 lambda: b"U" * 128, lambda uid, r, w: calls.append(("init", uid, r, w)))
assert calls[0] == ("init", b"U" * 128, rank, world), calls
nb = 1 - rank
assert calls[1] == (1 if rank == 0 else 0, b"window-of-%d" % nb), calls
assert D.reduce_max(float(rank + 1)) == 2.0
assert D.reduce_sum(float(rank + 1)) == 3.0
assert D.allgather_bytes(bytes([rank])) == [b"\x00", b"\x01"]
dist.destroy_process_group()
print("worker ok", rank)
'''


def test_two_rank_bootstrap_on_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MSP_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
                         env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("worker ok") == 2


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (CPU restatement on the host cores): one JSON line with the contract's keys."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0", "--cpu-sample-n", "256"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["higher_is_better"] is False and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"] and line["dtype"] == "f64"


def test_bench_gpu_arm_refuses_to_run_without_gpu():
    """No CPU fallback: on a box without CUDA the product arm of bench.py exits non-zero and says so."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_tsqr_root_is_exact_least_squares():
    """msp_tsqr_combine is host arithmetic (the root of the TSQR tree: Householder QR of the stacked (s+1)x(s+1) factors), so
    it is checked here without a GPU: factors made by numpy's QR of random row blocks [R_K | b_K] must give the least-squares
    solution and residual norm of the stacked system, for every basis size the drivers use and for 1..8 blocks."""
    import numpy as np
    from medane_tchakorom_ufc_thesis_repository_b200 import solver as S
    rng = np.random.default_rng(3)
    for s in (1, 3, 4, 5, 10, 20):
        for G in (1, 2, 4, 8):
            rows = 3 * (s + 1)
            blocks = [rng.standard_normal((rows, s + 1)) for _ in range(G)]
            factors = []
            for B in blocks:
                U = np.linalg.qr(B, mode="r")                       # (s+1) x (s+1) upper factor of [R_K | b_K]
                factors.append(np.asfortranarray(U).reshape(-1, order="F"))  # column-major, as op_local_qr hands it over
            alpha, rn = S.tsqr_combine(s, factors)
            A = np.vstack(blocks)
            ref, res, *_ = np.linalg.lstsq(A[:, :s], A[:, s], rcond=None)
            assert np.allclose(alpha, ref, rtol=1e-10, atol=1e-12), (s, G)
            assert abs(rn - np.sqrt(res[0])) <= 1e-10 * np.sqrt(res[0])
    # a numerically dependent basis vector (two identical iterates) is dropped, not divided by ~0
    B = rng.standard_normal((12, 4))
    B[:, 1] = B[:, 0]
    U = np.linalg.qr(B, mode="r")
    alpha, rn = S.tsqr_combine(3, [np.asfortranarray(U).reshape(-1, order="F")])
    assert np.all(np.isfinite(alpha)) and np.isfinite(rn)
    assert abs(np.linalg.norm(B[:, :3] @ alpha - B[:, 3]) - rn) <= 1e-8 * max(rn, 1.0)


def test_isolve_accepts_the_reference_option_spellings(tmp_path):
    """iSolve is the reference's launcher interface (reference iSolve:94-115, :60-200): its own option spellings must be
    accepted and end up, like there, as -inner_ksp_* / -outer_ksp_* options of the solver binary.  The binary is replaced by
    a stub that prints its arguments (MSOLVE=...), so nothing here needs a GPU."""
    stub = tmp_path / "msolve_stub"
    stub.write_text("#!/bin/sh\necho \"STUB $@\"\n")
    stub.chmod(0o755)
    env = dict(os.environ, MSOLVE=str(stub))
    isolve = os.path.join(ROOT, "iSolve")

    def run(*args):
        return subprocess.run([isolve, *args], capture_output=True, text=True, env=env, timeout=60)

    out = run("--alg", "SMSM_GLOBAL", "--np", "2", "--npb", "1", "--m", "512", "--n", "256", "--s", "5", "--rtol", "1e-6",
              "--inner-ksp", "gmres", "--inner-rtol", "1e-10", "--inner-max-iters", "20", "--inner-pc-type", "none",
              "--outer-ksp", "lsqr", "--outer-rtol", "1e-15", "--outer-max-iters", "70", "--outer-pc-type", "none",
              "--other-petsc-options", "-log_view -minimizer lsqr")
    assert out.returncode == 0, out.stderr
    assert "Program : Synchronous Multisplitting & Synchronous Minimization (GLOBAL MINIMIZATION)" in out.stdout
    assert "Mesh size : 512 x 256" in out.stdout and "OUTER solver max iterations : 70" in out.stdout
    assert "RUNNING COMMAND:" in out.stdout
    stub_line = [l for l in out.stdout.splitlines() if l.startswith("STUB ")][0].split()[1:]
    opts = {stub_line[i]: stub_line[i + 1] for i in range(len(stub_line) - 1) if stub_line[i].startswith("-")}
    assert opts["-alg"] == "SMSM_GLOBAL" and opts["-nblocks"] == "2" and opts["-m"] == "512" and opts["-n"] == "256"
    assert opts["-s"] == "5" and opts["-rtol"] == "1e-6"
    assert opts["-inner_ksp_type"] == "gmres" and opts["-inner_ksp_rtol"] == "1e-10" and opts["-inner_ksp_max_it"] == "20"
    assert opts["-inner_pc_type"] == "none"
    assert opts["-outer_ksp_type"] == "lsqr" and opts["-outer_ksp_rtol"] == "1e-15" and opts["-outer_ksp_max_it"] == "70"
    assert opts["-minimizer"] == "lsqr" and "-log_view" in stub_line
    assert "-min_convergence_count" not in opts                       # synchronous binaries do not get it (iSolve:362)
    # defaults of config/default_run_variables: AM, 2 processes, 1024 x 1024, s = 4, rtol 1e-3, inner gmres 20 / 1e-3
    out = run("--convergence_count_min", "7")
    stub_line = [l for l in out.stdout.splitlines() if l.startswith("STUB ")][0].split()[1:]
    opts = {stub_line[i]: stub_line[i + 1] for i in range(len(stub_line) - 1) if stub_line[i].startswith("-")}
    assert opts["-alg"] == "AM" and opts["-m"] == "1024" and opts["-n"] == "1024" and opts["-rtol"] == "1e-3"
    assert opts["-min_convergence_count"] == "7" and opts["-inner_ksp_max_it"] == "20" and "-s" not in opts
    assert "-outer_ksp_type" not in opts
    # np / npb that do not give whole blocks: the reference prints this and exits 0 (iSolve:332-338)
    out = run("--np", "3", "--npb", "2")
    assert out.returncode == 0 and "are not matching" in out.stdout and "STUB" not in out.stdout
    # several blocks (generalisation), stand-alone GMRES takes un-prefixed options, unknown things are refused
    out = run("--alg", "SM", "--np", "8", "--npb", "1")
    assert "-nblocks 8" in out.stdout
    out = run("--alg", "GMRES", "--inner-max-iters", "500")
    assert "-ksp_max_it 500" in out.stdout and "-inner_ksp_max_it" not in out.stdout and "-nblocks 1" in out.stdout
    assert run("--alg", "SMSM").returncode == 2
    assert run("--inner-ksp", "preonly").returncode == 2
    assert run("--bogus", "1").returncode == 2
    assert run("--m").returncode == 1
