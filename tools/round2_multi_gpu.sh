#!/bin/bash
# Second GPU call of the next round (after tools/round2_first_call.sh): the multi-rank parity check on N GPUs of one
# box -- the periodic bricks (verified in round 1) and, for the first time on devices, the laser-acceleration deck on
# slabs along z (golden checksums must not depend on the decomposition) -- then the scaling bench line.
# Usage:  gpurun --gpus 2 --timeout 1200 -- 'bash tools/round2_multi_gpu.sh 2'      (then 4)
set -u
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py > gpurun_out/multi_gpu_check_$N.txt 2>&1
echo "exit: $?" >> gpurun_out/multi_gpu_check_$N.txt
tail -40 gpurun_out/multi_gpu_check_$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 8 --warmup 4 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
cat gpurun_out/bench_$N.json
