#!/bin/bash
# Third GPU call of the next round: profiles of the chosen kernel variants (after tools/ab_modes.py picked them).
#   Usage:  gpurun --timeout 1500 -- 'bash tools/round2_profile.sh <deposit_mode> <gather_mode>'
# Brings back: the launch list of one bench run (shares of the step), and one `--set full` capture each of the
# deposition and the gather kernel (128^3 cells keep the ~40 replays per launch short).  Read here with
#   ncu -i gpurun_out/<name>.ncu-rep --page raw --csv | python profiles/summarize.py
#   ncu -i gpurun_out/<name>.ncu-rep --page source --csv | python profiles/srcstalls.py
set -u
D=${1:-0}
G=${2:-0}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_d${D}_g${G}.csv \
    python bench.py --steps 4 --warmup 3 --deposit-mode $D --gather-mode $G --profile-only > gpurun_out/launches_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:deposit_quiet -s 6 -c 2 -o gpurun_out/deposit_d${D} \
    python bench.py --cells 128 --steps 2 --warmup 3 --deposit-mode $D --gather-mode $G --profile-only > gpurun_out/ncu_deposit.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gather_push_ -s 6 -c 2 -o gpurun_out/gather_g${G} \
    python bench.py --cells 128 --steps 2 --warmup 3 --deposit-mode $D --gather-mode $G --profile-only > gpurun_out/ncu_gather.log 2>&1
ls -la gpurun_out | tail -8
