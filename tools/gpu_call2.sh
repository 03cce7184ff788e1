#!/bin/bash
# GPU session: parity tests incl. the rows around the path; ncu source-level capture of the backward tile kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_bwd_kernel -s 2 -c 2 -o gpurun_out/prof_bwd \
  python bench.py --no-graph --steps 1 --warmup 3 --skip-e2e --cpu-steps 0 > gpurun_out/ncu_bwd.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_bwd.log; ls -la gpurun_out/*.ncu-rep
