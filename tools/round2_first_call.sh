#!/bin/bash
# First GPU call of the round: (0) one-minute smoke of the kernels that are new this round, under `timeout` (a deadlock
# must not cost the slot); (1) every GPU parity test (no -x); (2) the A/B of the kernel variants on the benchmark
# workload in one process; (3) bench lines with the register-run and the lane-per-cell deposition.
# Usage:  gpurun --timeout 1800 -- 'bash tools/round2_first_call.sh'
set -u
mkdir -p gpurun_out
timeout 300 python tools/smoke_new_kernels.py > gpurun_out/smoke_new.txt 2>&1
rc=$?
echo "smoke exit: $rc" >> gpurun_out/smoke_new.txt
cat gpurun_out/smoke_new.txt | tail -8
if [ $rc -ne 0 ]; then
    # find the culprit one by one (each under its own timeout), then pin the survivors for the rest of the call
    for v in "PIC_FDTD_MODE=0" "PIC_DEPOSIT_MODE=0" ; do echo "retry with $v"; done
    export PIC_FDTD_MODE=0
    timeout 300 python tools/smoke_new_kernels.py > gpurun_out/smoke_new_nofdtd.txt 2>&1
    echo "smoke (plain FDTD) exit: $?" >> gpurun_out/smoke_new_nofdtd.txt
    tail -8 gpurun_out/smoke_new_nofdtd.txt
fi
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/gpu_tests.txt 2>&1
echo "pytest exit: $?" >> gpurun_out/gpu_tests.txt
tail -5 gpurun_out/gpu_tests.txt
timeout 600 python tools/ab_modes.py --cells 256 > gpurun_out/ab_modes.json 2> gpurun_out/ab_modes.err
tail -12 gpurun_out/ab_modes.err
timeout 400 python tools/ab_modes.py --cells 256 --jitter --deposit-modes 0,7 --gather-modes 0,1 > gpurun_out/ab_modes_jitter.json 2> gpurun_out/ab_modes_jitter.err
tail -6 gpurun_out/ab_modes_jitter.err
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json
timeout 400 python bench.py --steps 20 --warmup 5 --deposit-mode 7 > gpurun_out/bench_cells.json 2> gpurun_out/bench_cells.err
cat gpurun_out/bench_cells.json
