#!/bin/bash
mkdir -p gpurun_out
PROF_UPD_MODES=0,16,0 timeout 90 python tools/prof_tiles.py wd5m-upd 2>&1 | tail -6
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "prof failed or hung: stopping"; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x -k "touched or adagrad_in or large_table" > gpurun_out/pytest_upd.log 2>&1; rc=$?; echo "upd rc=$rc"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_upd.log | cut -c1-300 | head -20
[ $rc -ne 0 ] && exit 1
timeout 300 python bench_extra.py --workload wd5m-1vsall --batch 1024 --steps 10 --kernels > gpurun_out/extra_1vsall_upd.log 2>&1; grep '^{"workload' gpurun_out/extra_1vsall_upd.log | cut -c1-220; head -6 gpurun_out/kernels_wd5m_1vsall_B1024.txt | cut -c1-120
