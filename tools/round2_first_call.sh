#!/bin/bash
# First GPU call of the round: (1) every GPU parity test (no -x), (2) the A/B of the kernel variants on the benchmark
# workload in one process, (3) bench lines with the register-run and the lane-per-cell deposition.
# Usage:  gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/gpu_tests.txt 2>&1
echo "pytest exit: $?" >> gpurun_out/gpu_tests.txt
tail -5 gpurun_out/gpu_tests.txt
python tools/ab_modes.py --cells 256 > gpurun_out/ab_modes.json 2> gpurun_out/ab_modes.err
tail -12 gpurun_out/ab_modes.err
python tools/ab_modes.py --cells 256 --jitter --deposit-modes 0,7 --gather-modes 0,1 > gpurun_out/ab_modes_jitter.json 2> gpurun_out/ab_modes_jitter.err
tail -6 gpurun_out/ab_modes_jitter.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json
python bench.py --steps 20 --warmup 5 --deposit-mode 7 > gpurun_out/bench_cells.json 2> gpurun_out/bench_cells.err
cat gpurun_out/bench_cells.json
