"""CPU tests (-m "not gpu"): the warp-level deposition kernels of warpx_b200/csrc/deposit_runs.cu run
UNMODIFIED under the SIMT emulator of tests/host_harness/simt_host.h (every CUDA thread a cooperative
fiber, warp collectives as lock-step exchanges) against the oracle.  Covers the default register-run
kernel and the three experimental variants of pic_set_deposit_mode (two lines per lane, per-slot
reductions), cell-sorted, jittered (stale order: lone-particle and flush paths) and shuffled particles,
and the general kernel for the particles that changed cell.  The emulator checks indexing, lane roles and
arithmetic -- not memory-model races or speed."""
import ctypes as C

import numpy as np
import pytest

from helpers import lower_corner, rel_linf
from warpx_b200 import abi, workloads


@pytest.fixture(scope="module")
def simt():
    from host_harness import harness
    return harness.simt()


CASES = [(3, 0.02, "sorted"), (3, 0.5, "sorted"), (3, 0.02, "jitter"), (3, 0.3, "shuffled"), (1, 0.02, "jitter"),
         (2, 0.3, "sorted")]


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 6])
@pytest.mark.parametrize("nox,u_th,order", CASES)
def test_deposit_runs_kernels_under_simt_emulation(orc, simt, variant, nox, u_th, order):
    if variant and (nox, order) not in ((3, "sorted"), (3, "jitter"), (1, "jitter"), (2, "sorted")):
        pytest.skip("the experimental variants are covered on a subset")
    if variant >= 4 and nox != 3:
        pytest.skip("four lines per lane: order 3 only")
    n, lx = (8, 6, 6), (4e-6, 3e-6, 3e-6)
    wl = workloads.uniform_plasma_3d(n_cell=n, ppc=(2, 2, 2), u_th=u_th, lx=lx, seed=3)
    s = wl["species"][0]
    prob_lo = wl["prob_lo"]
    dx = [(wl["prob_hi"][d] - prob_lo[d]) / n[d] for d in range(3)]
    dinv = [1.0 / v for v in dx]
    dt = 0.95 / (np.sqrt(sum(1.0 / v ** 2 for v in dx)) * workloads.C)
    rng = np.random.default_rng(5)
    idx = np.arange(len(s["x"]))
    if order == "shuffled":
        rng.shuffle(idx)
    if order == "jitter":      # moved inside +-0.6 cell after the (cell-sorted) lattice order was fixed
        for d, k in enumerate("xyz"):
            s[k] = s[k] + rng.uniform(-0.6, 0.6, len(idx)) * dx[d]
    P = orc.HostParticles(**{k: s[k][idx] for k in orc.HostParticles.NAMES})
    ngJ = (nox + 2,) * 3
    xyzmin, lo = lower_corner(prob_lo, dx, (0, 0, 0), ngJ)
    box_hi = tuple(v - 1 for v in n)
    J = [orc.HostFab((0, 0, 0), box_hi, ngJ, abi.YEE_STAG[c]) for c in (6, 7, 8)]
    K = [orc.HostFab((0, 0, 0), box_hi, ngJ, abi.YEE_STAG[c]) for c in (6, 7, 8)]
    assert orc.lib().orc_deposit_esirkepov(C.byref(P.soa), 0, P.np, orc.fab_array(J), abi.dbl3(dinv), abi.dbl3(xyzmin),
                                           abi.int3(lo), s["q"], dt, -0.5 * dt, nox) == 0
    assert simt.simt_deposit_runs(C.byref(P.soa), 0, P.np, orc.fab_array(K), abi.dbl3(dinv), abi.dbl3(xyzmin),
                                  abi.int3(lo), s["q"], dt, -0.5 * dt, nox, variant) == 0
    for c in range(3):
        assert rel_linf(K[c].a, J[c].a) <= 1e-13, "j" + "xyz"[c]
