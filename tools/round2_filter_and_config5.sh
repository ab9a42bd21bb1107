#!/bin/bash
# 1-GPU call: the GPU suite (new filter tests included), the benchmark with the bilinear filter on (stage filter_j and its
# share of the HBM roofline), PIC_FILTER_DIRECT=1 for comparison, and BASELINE configs[4] (64 ppc stress) at 128^3 / 160^3.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/gpu_tests_r2k.txt 2>&1
echo "pytest exit: $?" >> gpurun_out/gpu_tests_r2k.txt
tail -3 gpurun_out/gpu_tests_r2k.txt
timeout 300 python bench.py --steps 20 --warmup 5 --filter 1 --profile-only > gpurun_out/bench_filter_on.json 2> gpurun_out/bench_filter_on.err
echo "filter bench exit: $?"
PIC_FILTER_DIRECT=1 timeout 300 python bench.py --steps 20 --warmup 5 --filter 1 --profile-only > gpurun_out/bench_filter_direct.json 2> gpurun_out/bench_filter_direct.err
echo "direct filter bench exit: $?"
timeout 400 python bench.py --cells 128 --ppc 4 --steps 20 --warmup 5 --cpu-cells 64 > gpurun_out/bench_config5_128.json 2> gpurun_out/bench_config5_128.err
echo "config 5 (128^3 x 64 ppc) exit: $?"
timeout 400 python bench.py --cells 160 --ppc 4 --steps 20 --warmup 5 --profile-only > gpurun_out/bench_config5_160.json 2> gpurun_out/bench_config5_160.err
echo "config 5 (160^3 x 64 ppc) exit: $?"
python - <<'PY'
import json
for f in ("bench_filter_on", "bench_filter_direct", "bench_config5_128", "bench_config5_160"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, round(d["ms_per_step"], 3), "%.3e" % d["value"], {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.05})
        if "filter_j" in d["roofline"]["kernels"]:
            print("   filter_j", d["roofline"]["kernels"]["filter_j"])
    except Exception as exc:
        print(f, "failed:", exc)
PY
