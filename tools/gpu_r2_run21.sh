#!/bin/bash
mkdir -p gpurun_out
PROF_UPD_MODES=0,24,16,7,3,4,8,32,0 timeout 600 python tools/prof_tiles.py wd5m-upd 2>&1 | tail -12
