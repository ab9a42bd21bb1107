#!/bin/bash
# Build the host-emulated library (tests/host_harness) with AddressSanitizer and run a python script against it:
#   tools/host_asan.sh script.py args...        (debugging aid; the CPU suite uses the plain build)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
HH=$ROOT/tests/host_harness
python -c "import sys; sys.path.insert(0, '$ROOT'); sys.path.insert(0, '$ROOT/tests'); from host_harness import harness; harness.build_host_library()"
OUT=$HH/_build/asan; mkdir -p $OUT
CUDA=$(dirname $(dirname ${NVCC:-/usr/local/cuda/bin/nvcc}))
for f in $HH/_build/host_src/*.cu.cpp; do
  o=$OUT/$(basename $f).o
  if [ ! -f $o ] || [ $f -nt $o ]; then
    g++ -std=c++17 -O1 -g -fno-omit-frame-pointer -fsanitize=address -fPIC -DPIC_SIMT_HOST -DPIC_HOST_HARNESS $PIC_ASAN_DEFS \
      -include $HH/simt_host.h -I $CUDA/include -I $HH/_build/host_src -ffp-contract=off -Wno-attributes \
      -Wno-unknown-pragmas -c $f -o $o &
  fi
done
wait
g++ -shared -fsanitize=address -o $OUT/libpic_host_asan.so $OUT/*.o -L $CUDA/lib64 -lcudart -ldl
export PIC_HOST_LIBRARY=$OUT/libpic_host_asan.so
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1
LD_PRELOAD=$(gcc -print-file-name=libasan.so) exec python "$@"
