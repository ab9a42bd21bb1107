#!/bin/bash
# Lean multi-GPU call (no CPU reference arm): the multi-rank parity check, then the scaling bench line as the driver
# launches it; with a second argument also the same bench with PIC_MIGRATE_FULL_SWEEP=1 (the round-1 classification).
# Usage:  gpurun --gpus N --timeout 900 -- 'bash tools/round2_multi_gpu_lean.sh N [ab]'
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py > gpurun_out/multi_gpu_check_$N.txt 2>&1
echo "exit: $?" >> gpurun_out/multi_gpu_check_$N.txt
grep -v "^\[rank[0-9]*\]:  \|^  File\|^    " gpurun_out/multi_gpu_check_$N.txt | tail -30 | cut -c1-220
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 --profile-only > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
echo "bench exit: $?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$N.json"))
print(d["ms_per_step"], d["value"], {k: round(v, 3) for k, v in d["stage_ms"].items()})
PY
if [ "${2:-}" = "ab" ]; then
    PIC_MIGRATE_FULL_SWEEP=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
        --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --profile-only > gpurun_out/bench_fullsweep_$N.json 2> gpurun_out/bench_fullsweep_$N.err
    echo "full-sweep bench exit: $?"
    python - <<PY
import json
d = json.load(open("gpurun_out/bench_fullsweep_$N.json"))
print(d["ms_per_step"], d["value"], {k: round(v, 3) for k, v in d["stage_ms"].items()})
PY
fi
