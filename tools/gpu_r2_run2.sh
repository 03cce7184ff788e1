#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_shapes.py -m gpu -q -rf > gpurun_out/pytest_run2.log 2>&1; echo "pytest rc=$?"
grep -v "^\[\|^  \|Warning\|^$" gpurun_out/pytest_run2.log | grep -n "^E \|FAILED\|passed\|failed\|Error" | head -80
