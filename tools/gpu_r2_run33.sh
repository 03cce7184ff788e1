#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_shapes.py tests/test_gpu_dropin.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_fl.log 2>&1; rc=$?; echo "rc=$rc"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_fl.log | cut -c1-400 | head -20
[ $rc -ne 0 ] && exit 1
bash tools/gpu_r2_2gpu.sh
