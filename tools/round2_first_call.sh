#!/bin/bash
# The first GPU call of the next round, in the order the minutes should go:
#   1. the GPU parity tests (several were written after round 1's GPU minutes ran out and have never run),
#   2. the A/B of the kernel variants on the benchmark workload (one process, ~2 min),
#   3. a bench line with the default modes.
# Usage:  gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/gpu_tests.txt 2>&1
echo "pytest exit: $?" >> gpurun_out/gpu_tests.txt
tail -5 gpurun_out/gpu_tests.txt
python tools/ab_modes.py --cells 256 > gpurun_out/ab_modes.json 2> gpurun_out/ab_modes.err
cat gpurun_out/ab_modes.err | tail -12
python bench.py --steps 8 --warmup 4 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json
