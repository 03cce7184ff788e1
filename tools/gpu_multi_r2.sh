#!/bin/bash
# usage: gpurun --gpus N --timeout 900 -- bash tools/gpu_multi_r2.sh N
# round 2: correctness of the sharded / replicated steps at N GPUs, then the bench line (Wikidata5M shape, row-sharded,
# strong scaling) with the exchanges as peer-memory kernels and, for comparison, over NCCL.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
tools/run_bounded.sh 240 gpurun_out/multi_gpu_check_$N.log $TR --master-port 29510 tests/multi_gpu_check.py; echo "multi_gpu_check rc=$?"
grep "==\|Error\|error\|assert" gpurun_out/multi_gpu_check_$N.log | head -20
tools/run_bounded.sh 200 gpurun_out/p2p_check_$N.log $TR --master-port 29511 tests/p2p_check.py; echo "p2p_check rc=$?"
grep "peer-memory\|Error\|error" gpurun_out/p2p_check_$N.log | head -5
tools/run_bounded.sh 400 gpurun_out/bench_wd5m_$N.log $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3; echo "bench rc=$?"
grep '^{"metric' gpurun_out/bench_wd5m_$N.log | cut -c1-400; grep -i "error\|unavailable" gpurun_out/bench_wd5m_$N.log | head -5
KGEB_NO_PEER=1 tools/run_bounded.sh 400 gpurun_out/bench_wd5m_${N}_nccl.log $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 --skip-e2e; echo "bench nccl rc=$?"
grep '^{"metric' gpurun_out/bench_wd5m_${N}_nccl.log | cut -c1-300
