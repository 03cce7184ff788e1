#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench_extra.py --workload wd5m-1vsall --batch 1024 --steps 40 --scale 0.125 --kernels > gpurun_out/extra_1vsall_eighth.log 2>&1; grep '^{"workload' gpurun_out/extra_1vsall_eighth.log | cut -c1-260; head -40 gpurun_out/kernels_wd5m_1vsall_B1024.txt | cut -c1-150
