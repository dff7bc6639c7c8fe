"""Micro-benchmark of the hot kernels on resident data: algorithmic GB/s per SURVEY.md §8(d).
usage: python tools/kbench.py [N | MxN] [restart] [3d]   (2-D N x N or M x N, or 3-D N^3; one block on cuda:0)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medane_tchakorom_ufc_thesis_repository_b200 import solver as S  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "4096"
M, N = (int(t) for t in shape.split("x")) if "x" in shape else (int(shape), int(shape))
restart = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dim3 = len(sys.argv) > 3 and sys.argv[3] == "3d"
peak = 6453.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
s_basis = int(os.environ.get("KBENCH_S", "5"))
if dim3:
    e = S.Engine(N, N, N, s=s_basis, max_restart=restart)
    W = 7
else:
    e = S.Engine(M, N, s=s_basis, max_restart=restart)
    W = 5
n = e.nb
nnz = (5 * M * N - 2 * M - 2 * N) if not dim3 else (7 * N ** 3 - 6 * N * N)
rows = []


def rec(name, ms, bytes_):
    gbs = bytes_ / (ms * 1e-3) / 1e9
    rows.append((name, ms, gbs, gbs / peak))
    print(f"{name:28s} {ms:9.4f} ms  {gbs:9.1f} GB/s  {gbs / peak:6.3f} of measured HBM peak ({peak:.0f} GB/s)", flush=True)


rec("copy (STREAM)", e.bench_kernel(4, iters=20), 16 * n)
fmt, width = e.spmv_format()
ms = e.bench_kernel(0, iters=20)
rec(f"spmv {fmt.upper()} [CSR bytes]", ms, 12 * nnz + 4 * (n + 1) + 16 * n)
rec(f"spmv {fmt.upper()} [own {width}-wide bytes]", ms, S.Engine.spmv_bytes_per_row(fmt, width) * n)
ms = e.bench_kernel(5, iters=20)
rec("spmv, input scaled on the fly", ms, 12 * nnz + 4 * (n + 1) + 16 * n)
for nv in (1, 2, 4, 8, 9, 12, 16, 17, 20, 24, 25, 30):
    if nv > restart:
        continue
    rec(f"mdot nv={nv}", e.bench_kernel(1, nv, iters=10), 8 * n * (nv + 1))
for nv in (1, 4, 8, 16, 30):
    if nv > restart:
        continue
    rec(f"maxpy+norm nv={nv}", e.bench_kernel(2, nv, iters=10), 8 * n * (nv + 2))
rec("gram [R|b]^T[R|b] 6 columns", e.bench_kernel(6, 6, iters=10), 8 * n * 6)
rec("spmm s=5", e.bench_kernel(3, 5, iters=10), 12 * nnz + 4 * n + 8 * n * 5 + 8 * n * 5)
if e.s >= 20:
    for nc in (11, 21):
        rec(f"gram, panels of 8, {nc} columns", e.bench_kernel(7, nc, iters=10), 8 * n * nc)
        rec(f"C := C T, panels of 8, {nc} columns", e.bench_kernel(8, nc, iters=10), 16 * n * nc)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"shape": shape, "dim": 3 if dim3 else 2, "spmv_format": fmt, "rows": rows},
          open(f"gpurun_out/kbench_{shape}{'_3d' if dim3 else ''}.json", "w"), indent=1)
