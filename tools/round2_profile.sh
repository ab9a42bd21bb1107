#!/bin/bash
# Profiling call: steady-state timings of the current variants, then one `ncu --set full` capture each of the
# lane-per-cell deposition, the supercell gather and the two bulk-staged FDTD kernels.
#   Usage:  gpurun --timeout 1500 -- 'bash tools/round2_profile.sh'
# Read here with   ncu -i gpurun_out/<name>.ncu-rep --page raw --csv | python profiles/summarize.py
#                  ncu -i gpurun_out/<name>.ncu-rep --page source --csv | python profiles/srcstalls.py
set -u
mkdir -p gpurun_out
timeout 500 python tools/ab_modes.py --cells 256 --jitter --deposit-modes 0,7 --gather-modes 0 > gpurun_out/ab2.json 2> gpurun_out/ab2.err
tail -8 gpurun_out/ab2.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:deposit_cells -s 4 -c 1 -f -o gpurun_out/r2_cells \
    python bench.py --cells 128 --spinup 0 --jitter --steps 2 --warmup 3 --deposit-mode 7 --profile-only > gpurun_out/ncu_cells.log 2>&1
timeout 300 $NCU -k regex:gather_push_tile -s 4 -c 1 -f -o gpurun_out/r2_gather \
    python bench.py --cells 128 --spinup 0 --jitter --steps 2 --warmup 3 --profile-only > gpurun_out/ncu_gather.log 2>&1
timeout 300 $NCU -k regex:evolve_._bulk -s 6 -c 2 -f -o gpurun_out/r2_fdtd \
    python bench.py --cells 256 --spinup 0 --steps 2 --warmup 3 --profile-only > gpurun_out/ncu_fdtd.log 2>&1
ls -la gpurun_out | tail -8
