"""Structural checks of the built library that need no GPU: the hot kernels are in libmsplit.so, compiled for sm_100a
only, within the register budgets their launch geometry assumes (a silent change there costs occupancy, not correctness),
and the cooperative restart-cycle kernel fits two blocks per SM."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "medane_tchakorom_ufc_thesis_repository_b200", "libmsplit.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not installed")


@pytest.fixture(scope="module")
def usage():
    assert os.path.exists(LIB), "libmsplit.so missing: run __graft_entry__.build()"
    out = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    res = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and name:
            res[name] = dict(reg=int(m.group(1)), stack=int(m.group(2)), shared=int(m.group(3)), local=int(m.group(4)))
            name = None
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    return res, archs


def _find(res, *needles):
    hits = [k for k in res if all(n in k for n in needles)]
    assert hits, f"no kernel matching {needles} in libmsplit.so"
    return hits


def test_sm_100a_only(usage):
    res, archs = usage
    assert archs == {"sm_100a"}, archs


def test_hot_kernels_present_and_within_register_budget(usage):
    res, _ = usage
    # coded-DIA stencil SpMV of the Arnoldi step (<ND, MODE 0, no residual, scaled input, no norm>): compiled for 4 blocks per SM
    for nd in (5, 7):
        for k in _find(res, "k_spmv_cdia_stencil", f"ILi{nd}ELi0ELb0ELb1ELb0E"):
            assert res[k]["reg"] <= 64 and res[k]["local"] == 0, (k, res[k])
    # VecMDot variants: the 24-vector one may take a whole SM's registers, the others two blocks per SM
    for k in _find(res, "k_mdot", "ILi24ELi1E"):
        assert res[k]["reg"] <= 255 and res[k]["local"] == 0, (k, res[k])
    for variant in ("ILi16ELi1E", "ILi8ELi2E", "ILi4ELi4E", "ILi2ELi8E"):
        for k in _find(res, "k_mdot", variant):
            assert res[k]["reg"] <= 128 and res[k]["local"] == 0, (k, res[k])
    for k in _find(res, "k_maxpy_norm"):
        assert res[k]["reg"] <= 128 and res[k]["local"] == 0, (k, res[k])
    for k in _find(res, "k_gram"):
        assert res[k]["local"] == 0, (k, res[k])


def test_cycle_kernel_fits_two_blocks_per_sm(usage):
    res, _ = usage
    for nd in (5, 7):
        (k,) = _find(res, "k_gmres_cycle_coop", f"ILi{nd}E")
        # 2 x 256 threads x 128 registers = the register file of an SM; static shared memory under the 48 KB static limit, twice
        # under the 227 KB of an SM
        assert res[k]["reg"] <= 128, (k, res[k])
        assert res[k]["shared"] <= 48 * 1024, (k, res[k])
        assert res[k]["stack"] <= 256, (k, res[k])
