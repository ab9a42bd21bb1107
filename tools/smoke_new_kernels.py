#!/usr/bin/env python3
"""Device smoke of the kernels that are new in this round, meant to run under `timeout` BEFORE the test suite (a kernel
that deadlocks must cost one minute, not the GPU slot):
  * fdtd_bulk.cu (cp.async.bulk + mbarrier ring) against the plain-load kernels of fdtd.cu,
  * deposit_cells.cu (PIC_DEPOSIT_CELLS) against the register-run kernels, sorted / stale bins.
Prints one line per check; exit status 0 only if all agree."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from warpx_b200 import abi, workloads          # noqa: E402
from warpx_b200.engine import Simulation      # noqa: E402
from warpx_b200.lib import lib                # noqa: E402


def run(n, steps, **modes):
    L = lib()
    L.pic_set_fdtd_mode(modes.get("fdtd", 1))
    L.pic_set_deposit_mode(modes.get("deposit", 0))
    L.pic_set_gather_mode(modes.get("gather", 0))
    wl = workloads.uniform_plasma_3d(n=n, ppc=(2, 2, 2), u_th=modes.get("u_th", 0.05), lx=40.0e-6 * n / 256.0, perturbation=0.01)
    sim = Simulation(wl["n_cell"], wl["prob_lo"], wl["prob_hi"], nox=3, sort_interval=4)
    s = wl["species"][0]
    sim.add_species(s["name"], s["q"], s["m"], s["x"], s["y"], s["z"], s["w"], s["ux"], s["uy"], s["uz"])
    sim.Evolve(steps)
    torch.cuda.synchronize()
    out = [sim.field_numpy(c)[1].copy() for c in range(9)]
    sim.close()
    L.pic_set_fdtd_mode(1); L.pic_set_deposit_mode(0); L.pic_set_gather_mode(0)
    return out


def main():
    ok = True
    torch.cuda.set_device(0)
    ref = run(32, 7, fdtd=0, deposit=0)
    for name, modes in (("fdtd bulk staging", dict(fdtd=3, deposit=0)), ("deposit lane-per-cell", dict(fdtd=0, deposit=7)),
                        ("deposit lane-per-cell, 2 producers", dict(fdtd=0, deposit=8)), ("the same, wide", dict(fdtd=0, deposit=9)),
                        ("decoupled pipeline", dict(fdtd=0, deposit=10)), ("decoupled pipeline, wide", dict(fdtd=0, deposit=11)),
                        ("gather pairs", dict(fdtd=0, deposit=0, gather=1)), ("all three", dict(fdtd=3, deposit=7, gather=1))):
        got = run(32, 7, **modes)
        err = max(float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) for a, b in zip(got[:3] + got[6:], ref[:3] + ref[6:]))
        good = err <= 1e-9
        ok &= good
        print("[smoke] %-36s rel. difference to the round-1 kernels %.2e  %s" % (name, err, "ok" if good else "FAIL"), flush=True)
    print("SMOKE_NEW_KERNELS", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
