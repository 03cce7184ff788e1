#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x -k "touched or adagrad_in or large_table" > gpurun_out/pytest_upd.log 2>&1; echo "upd rc=$?"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_upd.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_wd_upd.json 2> gpurun_out/bench_wd_upd.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_wd_upd.json; tail -3 gpurun_out/bench_wd_upd.err
timeout 600 python bench_extra.py --workload wd5m-1vsall --batch 1024 --steps 10 --kernels > gpurun_out/extra_1vsall_upd.log 2>&1; grep '^{"workload' gpurun_out/extra_1vsall_upd.log | cut -c1-400; head -12 gpurun_out/kernels_wd5m_1vsall_B1024.txt | cut -c1-170
