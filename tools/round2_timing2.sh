#!/bin/bash
# How much of the steady-state penalty is staleness of the bins, how much the Poisson counts: freshly sorted random
# positions; then the cell-sort interval sweep.
set -u
mkdir -p gpurun_out
timeout 400 python tools/ab_modes.py --cells 256 --jitter --fresh --deposit-modes 0 --gather-modes 0 > gpurun_out/ab5_fresh_jitter.json 2> gpurun_out/ab5_fresh_jitter.err
tail -3 gpurun_out/ab5_fresh_jitter.err
timeout 600 python tools/ab_modes.py --cells 256 --sort-intervals 1,2,3,4,6,8 > gpurun_out/ab5_sort.json 2> gpurun_out/ab5_sort.err
tail -8 gpurun_out/ab5_sort.err
