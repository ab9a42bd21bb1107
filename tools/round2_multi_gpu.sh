#!/bin/bash
# Multi-GPU call: the multi-rank parity check on N GPUs of one box (periodic bricks against WarpX's golden checksums and
# the single-box oracle; the laser-acceleration deck on slabs along z), then the scaling bench line exactly as the
# driver launches it (the FULL bench: timed steps, e2e leg, collective teardown, CPU leg on rank 0).
# Usage:  gpurun --gpus N --timeout 1200 -- 'bash tools/round2_multi_gpu.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py > gpurun_out/multi_gpu_check_$N.txt 2>&1
echo "exit: $?" >> gpurun_out/multi_gpu_check_$N.txt
tail -45 gpurun_out/multi_gpu_check_$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
echo "bench exit: $?"
cat gpurun_out/bench_$N.json
tail -5 gpurun_out/bench_$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_ref_$N.json 2> gpurun_out/bench_ref_$N.err
echo "reference arm exit: $?"
cat gpurun_out/bench_ref_$N.json
if [ "$N" = "8" ]; then
    # BASELINE configs[2]: 512^3 split 2x2x2 over 8 GPUs, CKC solver, NCCL halos
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
        bench.py --gpus $N --steps 20 --warmup 5 --solver ckc > gpurun_out/bench_ckc_$N.json 2> gpurun_out/bench_ckc_$N.err
    echo "ckc bench exit: $?"
    cat gpurun_out/bench_ckc_$N.json
fi
