#!/bin/bash
mkdir -p gpurun_out
PROF_UPD_MODES=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_bwd4_kernel -s 6 -c 1 -o gpurun_out/prof_upd python tools/prof_tiles.py wd5m-upd > gpurun_out/ncu_upd.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_upd.log
