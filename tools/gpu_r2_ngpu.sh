#!/bin/bash
# usage: gpurun --gpus N --timeout 900 -- bash tools/gpu_r2_ngpu.sh N [p2p]
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
tools/run_bounded.sh 200 gpurun_out/multi_gpu_check_$N.log $TR --master-port 29510 tests/multi_gpu_check.py; echo "multi_gpu_check $N rc=$?"
grep "==\|Error\|assert\|unavailable" gpurun_out/multi_gpu_check_$N.log | cut -c1-200 | head -16
if [ "$2" == "p2p" ]; then
  tools/run_bounded.sh 150 gpurun_out/p2p_check_$N.log $TR --master-port 29530 tests/p2p_check.py; echo "p2p_check $N rc=$?"
  grep "peer-memory\|Error\|assert" gpurun_out/p2p_check_$N.log | cut -c1-200 | head -5
fi
tools/run_bounded.sh 300 gpurun_out/bench_wd5m_$N.log $TR --master-port 29540 bench.py --gpus $N --steps 30 --warmup 3; echo "bench $N rc=$?"
grep '^{"metric' gpurun_out/bench_wd5m_$N.log | cut -c1-330; grep -i "error\|unavailable" gpurun_out/bench_wd5m_$N.log | head -3
