#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/smoke_new_kernels.py > gpurun_out/smoke_new.txt 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke_new.txt
tail -3 gpurun_out/smoke_new.txt
timeout 500 python tools/ab_modes.py --cells 256 --jitter --deposit-modes 0 --gather-modes 0 > gpurun_out/ab7.json 2> gpurun_out/ab7.err
tail -5 gpurun_out/ab7.err
