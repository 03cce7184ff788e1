#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/debug_flash.py 2>&1 | tail -10
echo "== flash + golden parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -q > gpurun_out/pytest_r6.log 2>&1; echo "rc=$?"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_r6.log | cut -c1-250 | head -12
timeout 200 python tools/prof_tiles.py wd5m > gpurun_out/prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_bwd4 -s 2 -c 2 -o gpurun_out/prof_wd5m_v2 python tools/prof_tiles.py wd5m > gpurun_out/ncu_wd5m.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_wd5m.log
