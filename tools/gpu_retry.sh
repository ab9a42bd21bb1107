#!/bin/bash
# Keeps asking for a GPU slot until the call actually runs (gpurun answers "transient" while the pod is busy).
#   tools/gpu_retry.sh <log> <gpurun args...>
log=$1; shift
for attempt in $(seq 1 40); do
    gpurun "$@" > "$log" 2>&1
    if ! grep -q "status=transient" "$log"; then exit 0; fi
    sleep 120
done
