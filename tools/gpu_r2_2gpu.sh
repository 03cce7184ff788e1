#!/bin/bash
# usage: gpurun --gpus 2 --timeout 700 -- bash tools/gpu_r2_2gpu.sh
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
tools/run_bounded.sh 200 gpurun_out/multi_gpu_check_$N.log $TR --master-port 29510 tests/multi_gpu_check.py; rc=$?; echo "multi_gpu_check rc=$rc"
grep "==\|Error\|error\|assert" gpurun_out/multi_gpu_check_$N.log | cut -c1-200 | head -20
[ $rc -ne 0 ] && exit 1
tools/run_bounded.sh 300 gpurun_out/bench_wd5m_$N.log $TR --master-port 29513 bench.py --gpus $N --steps 30 --warmup 3; echo "bench rc=$?"
grep '^{"metric' gpurun_out/bench_wd5m_$N.log | cut -c1-330; grep -i "error\|unavailable" gpurun_out/bench_wd5m_$N.log | head -5
