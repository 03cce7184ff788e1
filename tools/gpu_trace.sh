#!/bin/bash
# usage: tools/build_variant.sh trace tc_bwd -DKGEB_TRACE && gpurun -- bash tools/gpu_trace.sh   (pipeline timeline of the backward tile kernels)
mkdir -p gpurun_out
KGEB200_LIB=$PWD/kge-1_b200/variants/libkgeb200_trace.so timeout 300 python tools/trace_bwd.py bce > gpurun_out/trace_bce.txt 2>&1
KGEB200_LIB=$PWD/kge-1_b200/variants/libkgeb200_trace.so timeout 300 python tools/trace_bwd.py kl > gpurun_out/trace_kl.txt 2>&1
grep -v "^   \|^---" gpurun_out/trace_bce.txt | head -30
