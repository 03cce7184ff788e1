#!/bin/bash
# Builds libkgeb200.so (the C-ABI library, include/kgeb200.h) for sm_100a, in-tree.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libkgeb200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2"
mkdir -p "$HERE/build"
pids=()
for f in rows tiles update batch sampler penalty metrics p2p ns_segment tc_dot tc_bwd; do
  [ -f "$HERE/$f.cu" ] || continue
  if [ ! -f "$HERE/build/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/build/$f.o" ] || [ "$HERE/common.cuh" -nt "$HERE/build/$f.o" ] || [ "$HERE/tc_common.cuh" -nt "$HERE/build/$f.o" ] || [ "$HERE/../../include/kgeb200.h" -nt "$HERE/build/$f.o" ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c "$HERE/$f.cu" -o "$HERE/build/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o "$OUT" "$HERE"/build/*.o -lcudart
echo "built $OUT"
