#!/bin/bash
# GPU session: instruction-rate microbenchmark, parity tests, default bench, epilogue variants
mkdir -p gpurun_out
tools/bin/mufu_bench > gpurun_out/mufu_bench.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
for v in lg8 lg32 rp rplg8 rplg32; do
  KGEB200_LIB=$PWD/kge-1_b200/variants/libkgeb200_$v.so timeout 200 python bench.py --steps 30 --warmup 5 --cpu-steps 0 --skip-e2e > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
done
KGEB200_LIB=$PWD/kge-1_b200/variants/libkgeb200_rplg8.so timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_rplg8.log 2>&1
tail -3 gpurun_out/pytest_gpu.log gpurun_out/pytest_gpu_rplg8.log; cat gpurun_out/mufu_bench.txt
