#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x -k "adagrad_in" > gpurun_out/pytest_last.log 2>&1; rc=$?; echo "rc=$rc"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_last.log | cut -c1-600 | head -12
