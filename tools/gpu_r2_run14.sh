#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench_extra.py --workload wd5m-1vsall --batch 1024 --graph-step --kernels --steps 20 > gpurun_out/extra_1vsall_r14.log 2>&1; grep '^{"workload' gpurun_out/extra_1vsall_r14.log | cut -c1-200; head -40 gpurun_out/kernels_wd5m_1vsall_B1024.txt | cut -c1-170
