#!/bin/bash
# Timing call: device smoke of the new kernels (under timeout), then the A/B of the deposition variants in the steady
# state and on the fresh lattice.   Usage: gpurun --timeout 1200 -- 'bash tools/round2_timing.sh [modes]'
set -u
mkdir -p gpurun_out
timeout 300 python tools/smoke_new_kernels.py > gpurun_out/smoke_new.txt 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke_new.txt
tail -11 gpurun_out/smoke_new.txt
timeout 500 python tools/ab_modes.py --cells 256 --jitter --deposit-modes ${1:-0,7,10,11} --gather-modes 0 > gpurun_out/ab4.json 2> gpurun_out/ab4.err
tail -9 gpurun_out/ab4.err
timeout 500 python tools/ab_modes.py --cells 256 --fresh --deposit-modes ${1:-0,7,10,11} --gather-modes 0 > gpurun_out/ab4_fresh.json 2> gpurun_out/ab4_fresh.err
tail -6 gpurun_out/ab4_fresh.err
