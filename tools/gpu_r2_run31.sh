#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x -k "failed_flash_pass" > gpurun_out/pytest_fb.log 2>&1; rc=$?; echo "rc=$rc"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_fb.log | cut -c1-400 | head -20
