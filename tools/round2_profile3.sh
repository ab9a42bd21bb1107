#!/bin/bash
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:deposit_quiet\|deposit_general -s 8 -c 2 -f -o gpurun_out/r2_runs \
    python bench.py --cells 128 --spinup 0 --jitter --steps 2 --warmup 3 --profile-only > gpurun_out/ncu_runs.log 2>&1
tail -2 gpurun_out/ncu_runs.log | cut -c1-200
ls -la gpurun_out/r2_runs.ncu-rep
