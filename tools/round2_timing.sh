#!/bin/bash
# Timing call: device smoke of the new kernels (under timeout), then the steady-state A/B of the deposition variants and
# the FDTD data paths.   Usage: gpurun --timeout 1200 -- 'bash tools/round2_timing.sh'
set -u
mkdir -p gpurun_out
timeout 300 python tools/smoke_new_kernels.py > gpurun_out/smoke_new.txt 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke_new.txt
tail -9 gpurun_out/smoke_new.txt
timeout 700 python tools/ab_modes.py --cells 256 --jitter --deposit-modes ${1:-0,7,8,9} --gather-modes 0 > gpurun_out/ab3.json 2> gpurun_out/ab3.err
tail -10 gpurun_out/ab3.err
