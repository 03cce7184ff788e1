#!/bin/bash
# usage: gpurun --gpus 8 --timeout 1200 -- bash tools/gpu_r2_8gpu.sh
mkdir -p gpurun_out
tr() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1"; }
for N in 8 4; do
  tools/run_bounded.sh 200 gpurun_out/multi_gpu_check_$N.log $(tr $N) --master-port 2951$N tests/multi_gpu_check.py; echo "multi_gpu_check $N rc=$?"
  grep "==\|Error\|assert\|unavailable" gpurun_out/multi_gpu_check_$N.log | cut -c1-200 | head -16
  tools/run_bounded.sh 150 gpurun_out/ns_sharded_check_$N.log $(tr $N) --master-port 2952$N tests/ns_sharded_check.py; echo "ns_sharded_check $N rc=$?"
  grep "==\|Error\|assert" gpurun_out/ns_sharded_check_$N.log | cut -c1-200 | head -8
  tools/run_bounded.sh 150 gpurun_out/p2p_check_$N.log $(tr $N) --master-port 2953$N tests/p2p_check.py; echo "p2p_check $N rc=$?"
  grep "peer-memory\|Error\|assert" gpurun_out/p2p_check_$N.log | cut -c1-200 | head -5
  tools/run_bounded.sh 300 gpurun_out/bench_wd5m_$N.log $(tr $N) --master-port 2954$N bench.py --gpus $N --steps 30 --warmup 3; echo "bench $N rc=$?"
  grep '^{"metric' gpurun_out/bench_wd5m_$N.log | cut -c1-330; grep -i "error\|unavailable" gpurun_out/bench_wd5m_$N.log | head -3
done
nvidia-smi topo -m 2>/dev/null | head -12
