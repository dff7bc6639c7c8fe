#!/usr/bin/env python
"""Parameter sweep of the reference's shipped option space on the device engine.

The reference's bulk scripts run inner `-ksp_max_it` in {2,3,5,10,20,30,50}, `-s` in {4,5,10,20}, GMRES restart 30
(running_bulk_test_local:72-310, running_bulk_test_g5k:230-320).  This tool runs one algorithm over a grid of those
settings, every point capped by wall-clock seconds and by outer iterations, and appends one JSON line per point to
the output file as soon as the point ends (a cut-off call keeps what it finished).

All blocks can share ONE GPU (`--devices 0`, the default: the synchronous algorithms are deterministic in the block
count, not in the GPU count) or be spread over the GPUs of the box (`--devices all`).

  python tools/sweep.py --grid 128 --blocks 8 --algs SMSM_GLOBAL,SMSM_SEMI_LOCAL --max-it 20,30,50,60 \
      --restart 30,50 --s 4,5,10,20 --cap-seconds 15 --out gpurun_out/sweep_128.jsonl
"""
from __future__ import annotations

import argparse
import itertools
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=128, help="edge N of the N^3 (or N^2 with --dim 2) grid")
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--blocks", type=int, default=8)
    ap.add_argument("--algs", default="SMSM_GLOBAL")
    ap.add_argument("--max-it", default="20")
    ap.add_argument("--restart", default="30")
    ap.add_argument("--s", default="5")
    ap.add_argument("--refine", default="0")
    ap.add_argument("--minimizer", default="tsqr")
    ap.add_argument("--inner-rtol", default="1e-10")
    ap.add_argument("--rtol", type=float, default=1e-6)
    ap.add_argument("--cap-seconds", type=float, default=15.0)
    ap.add_argument("--cap-outer", type=int, default=100000)
    ap.add_argument("--devices", default="0")
    ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
    ap.add_argument("--budget-seconds", type=float, default=1e9, help="stop starting new points after this much wall time")
    args = ap.parse_args()

    import numpy as np  # noqa: F401
    from medane_tchakorom_ufc_thesis_repository_b200 import _lib
    from medane_tchakorom_ufc_thesis_repository_b200 import solver as S

    ngpu = _lib.lib().msp_device_count()
    if ngpu < 1:
        raise SystemExit("sweep.py needs a CUDA device")
    G = args.blocks
    devices = [k % ngpu for k in range(G)] if args.devices == "all" else [int(d) for d in args.devices.split(",")]
    devices = [devices[k % len(devices)] for k in range(G)]
    N = args.grid
    m, n, p = (N, N, N) if args.dim == 3 else (N, N, 1)
    ints = lambda t: [int(v) for v in t.split(",")]  # noqa: E731
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    t_begin = time.time()
    points = list(itertools.product(args.algs.split(","), ints(args.s), ints(args.restart), ints(args.max_it), ints(args.refine),
                                    args.minimizer.split(","), [float(v) for v in args.inner_rtol.split(",")]))
    cache = {}
    for alg, s, restart, max_it, refine, mini, irtol in points:
        if max_it <= 30 and restart > 30 and 30 in ints(args.restart):
            continue  # a restart longer than max_it changes nothing
        if time.time() - t_begin > args.budget_seconds:
            break
        key = (s, restart)
        if key not in cache:
            for g in cache.values():
                g.close()
            cache.clear()
            cache[key] = S.Group(m, n, p, nblocks=G, s=s, max_restart=restart, devices=devices)
        grp = cache[key]
        for e in grp.engines:  # x0 = 0, halos = 0: a fresh run on the resident matrix
            e.x = np.zeros(e.nb)
            for side in (0, 1):
                e.set_halo(side, np.zeros(e.H))
        inner = S.ksp_opts(restart=restart, max_it=max_it, rtol=irtol, abstol=1e-100, cgs_refine=refine)
        t0 = time.time()
        try:
            res = grp.solve(alg, s=s, rtol=args.rtol, inner=inner, max_outer=args.cap_outer, outer_type=mini,
                            max_seconds=args.cap_seconds)
            r0 = res[0]
            h = r0["hist"] / r0["norm0"]
            line = {
                "alg": alg, "grid": [m, n, p], "blocks": G, "s": s, "restart": restart, "inner_max_it": max_it, "inner_rtol": irtol,
                "cgs_refine": refine, "minimizer": mini, "rtol": args.rtol,
                "outer_its": r0["outer_its"], "inner_its_total_block0": int(r0["inner_its_total"]),
                "stop_reason": {0: "converged", 1: "max_outer", 2: "max_seconds"}[r0["stop_reason"]],
                "stopping_quantity_rel": r0["last_norm"] / r0["norm0"],
                "true_rel_residual": r0["final_residual"] / r0["norm0"],
                "reached_1e-6_true": bool(r0["final_residual"] <= args.rtol * r0["norm0"] * 1.000001),
                "device_s": max(r["elapsed_s"] for r in res), "wall_s": time.time() - t0,
                "hist_rel_sampled": [float(v) for v in h[:: max(1, len(h) // 16)]],
            }
        except Exception as ex:  # keep sweeping
            line = {"alg": alg, "grid": [m, n, p], "blocks": G, "s": s, "restart": restart, "inner_max_it": max_it,
                    "cgs_refine": refine, "minimizer": mini, "error": str(ex)}
        with open(args.out, "a") as f:
            f.write(json.dumps(line) + "\n")
        print(json.dumps({k: v for k, v in line.items() if k != "hist_rel_sampled"}), flush=True)
    for g in cache.values():
        g.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
