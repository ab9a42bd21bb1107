#!/bin/bash
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:deposit_cells2 -s 4 -c 1 -f -o gpurun_out/r2_cells2 \
    python bench.py --cells 128 --spinup 0 --jitter --steps 2 --warmup 3 --deposit-mode ${1:-8} --profile-only > gpurun_out/ncu_cells2.log 2>&1
tail -3 gpurun_out/ncu_cells2.log
ls -la gpurun_out/r2_cells2.ncu-rep
