#!/bin/bash
# A/B of the mover split in the default deposition (mode 0 = split, mode 12 = round-1 behaviour), steady state, after a
# device check of the split against the unsplit kernel.
set -u
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/smoke_split.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import numpy as np, torch
import smoke_new_kernels as S
ref = S.run(32, 7, fdtd=0, deposit=12, u_th=0.2)
got = S.run(32, 7, fdtd=0, deposit=0, u_th=0.2)
err = max(float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) for a, b in zip(got[:3] + got[6:], ref[:3] + ref[6:]))
print("[smoke] mover split vs general kernel, u_th = 0.2 c: rel. difference %.2e %s" % (err, "ok" if err <= 1e-9 else "FAIL"))
PY
cat gpurun_out/smoke_split.txt | tail -2
timeout 500 python tools/ab_modes.py --cells 256 --jitter --deposit-modes 0,12 --gather-modes 0 > gpurun_out/ab6.json 2> gpurun_out/ab6.err
tail -6 gpurun_out/ab6.err
