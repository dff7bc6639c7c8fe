"""Timing probe of the persistent cooperative restart-cycle kernel (csrc/cycle_coop.cuh) against one kernel per phase:
microseconds per Arnoldi step over block sizes, one engine alone and two engines sharing the GPU, then BASELINE
config 1 (MSM, 512 x 512, 2 blocks, inner GMRES(30) max_it 50 rtol 1e-10, rtol 1e-6) both ways.  Prints JSON lines."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medane_tchakorom_ufc_thesis_repository_b200 import solver as S  # noqa: E402

SIZES = [(128, 128, 1), (362, 364, 1), (512, 512, 1), (724, 724, 1), (1024, 1024, 1), (1448, 1448, 1), (2048, 2048, 1), (64, 64, 64), (128, 128, 128)]
if os.environ.get("PROBE_SIZES"):
    SIZES = [tuple(int(v) for v in t.split("x")) for t in os.environ["PROBE_SIZES"].split(",")]
MODES = os.environ.get("PROBE_MODES", "0,1").split(",")
CTAS = os.environ.get("PROBE_CTAS", "1,2").split(",")


def set_mode(mode, ctas):
    """mode "0": one kernel per phase; "1": the cycle kernel, forced, `ctas` blocks per SM; "auto": the library's default policy."""
    for k in ("MSPLIT_COOP", "MSPLIT_COOP_CTAS_PER_SM"):
        os.environ.pop(k, None)
    if mode != "auto":
        os.environ["MSPLIT_COOP"] = mode
        os.environ["MSPLIT_COOP_CTAS_PER_SM"] = ctas


def per_step(m, n, p, G, mode, ctas):
    set_mode(mode, ctas)
    grp = S.Group(m * G if p == 1 else m, n, p if p == 1 else p * G, nblocks=G, s=0, max_restart=30)
    rows = grp.engines[0].nb
    inner = S.ksp_opts(restart=30, max_it=30, rtol=1e-30, abstol=1e-300)
    grp.solve("SM", rtol=1e-300, inner=inner, max_outer=3)
    outer = 20 if rows < 3_000_000 else 6
    res = grp.solve("SM", rtol=1e-300, inner=inner, max_outer=outer)
    its = res[0]["inner_its_total"]
    out = dict(case="per_step", grid=f"{m}x{n}x{p}", rows_per_block=rows, blocks_on_gpu=G, mode=mode, ctas_per_sm=(int(ctas) if mode == "1" else None),
               persistent=bool(grp.engines[0].persistent_cycles()), us_per_arnoldi_step=res[0]["elapsed_s"] * 1e6 / max(its, 1),
               inner_its=its, launches=res[0]["kernel_launches"], rel_residual=res[0]["final_residual"] / res[0]["norm0"])
    grp.close()
    return out


def config1(mode, ctas):
    set_mode(mode, ctas)
    grp = S.Group(512, 512, nblocks=2, s=0, max_restart=30)
    inner = S.ksp_opts(restart=30, max_it=50, rtol=1e-10, abstol=1e-100)
    grp.solve("SM", rtol=1e-2, inner=inner, max_outer=5)
    for e in grp.engines:
        import numpy as np
        e.x = np.zeros(e.nb)
        for side in (0, 1):
            e.set_halo(side, np.zeros(e.H))
    res = grp.solve("SM", rtol=1e-6, inner=inner, max_outer=100000)
    out = dict(case="config1_msm_512_2blocks_1gpu", mode=mode, ctas_per_sm=(int(ctas) if mode == "1" else None), persistent=bool(grp.engines[0].persistent_cycles()),
               seconds=res[0]["elapsed_s"], outer_its=res[0]["outer_its"], inner_its=res[0]["inner_its_total"], launches=res[0]["kernel_launches"],
               rel_residual=res[0]["final_residual"] / res[0]["norm0"], stop_reason=res[0]["stop_reason"])
    grp.close()
    return out


if __name__ == "__main__":
    what = os.environ.get("PROBE_WHAT", "steps,config1").split(",")
    if "config1" in what:
        for mode in MODES:
            for ctas in (CTAS if mode == "1" else ["1"]):
                print(json.dumps(config1(mode, ctas)), flush=True)
    if "steps" in what:
        for (m, n, p) in SIZES:
            for G in (1, 2):
                for mode in MODES:
                    for ctas in (CTAS if mode == "1" else ["1"]):
                        print(json.dumps(per_step(m, n, p, G, mode, ctas)), flush=True)
