#!/bin/bash
# Final single-GPU call of the round: the whole GPU test suite (no -x), smoke(), the bench line as the driver runs it,
# the reference arm, and the ncu launch list of the bench command (shares of the step).
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/gpu_tests_final.txt 2>&1
echo "pytest exit: $?" >> gpurun_out/gpu_tests_final.txt
tail -4 gpurun_out/gpu_tests_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.txt 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke_final.txt
tail -2 gpurun_out/smoke_final.txt
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit: $?"; cat gpurun_out/bench_final.json | cut -c1-400
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
echo "reference exit: $?"; cat gpurun_out/bench_ref_final.json | cut -c1-300
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 4 --warmup 3 --spinup 12 --profile-only > gpurun_out/launches_run.log 2>&1
echo "ncu exit: $?"; tail -2 gpurun_out/launches_run.log | cut -c1-200
