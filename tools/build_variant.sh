#!/bin/bash
# Tuning helper: builds a variant of the C-ABI library with extra nvcc defines for ONE source file, e.g.
#   tools/build_variant.sh p53 tc_bwd "-DKGEB_POLY8_BCE=5 -DKGEB_POLY8_KL=3"
# -> kge-1_b200/variants/libkgeb200_p53.so (git-ignored; selected at run time with KGEB200_LIB=<path>).
set -e
NAME=$1; SRC=$2; DEFS=$3
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
C="$ROOT/kge-1_b200/csrc"
mkdir -p "$ROOT/kge-1_b200/variants" "$C/build/var_$NAME"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2"
$NVCC $FLAGS $DEFS -c "$C/$SRC.cu" -o "$C/build/var_$NAME/$SRC.o"
OBJS=""
for f in rows tiles update batch sampler penalty metrics p2p ns_segment tc_dot tc_bwd; do
  if [ "$f" == "$SRC" ]; then OBJS="$OBJS $C/build/var_$NAME/$f.o"; else OBJS="$OBJS $C/build/$f.o"; fi
done
$NVCC -shared -o "$ROOT/kge-1_b200/variants/libkgeb200_$NAME.so" $OBJS -lcudart
echo "built variants/libkgeb200_$NAME.so"
