#!/bin/bash
# usage: gpurun --gpus N --timeout 800 -- bash tools/gpu_multi.sh N   (peer-memory mode against NCCL: check + bench in both modes)
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
tools/run_bounded.sh 200 gpurun_out/p2p_check_$N.log $TR --master-port 29511 tests/p2p_check.py; echo "p2p_check rc=$?"
grep "peer-memory\|Error\|error" gpurun_out/p2p_check_$N.log | head -5
tools/run_bounded.sh 200 gpurun_out/bench${N}_p2p.log $TR --master-port 29513 bench.py --gpus $N --steps 50 --warmup 5 --skip-e2e --cpu-steps 0 --parallel p2p; echo "p2p rc=$?"
grep '^{"metric' gpurun_out/bench${N}_p2p.log | cut -c1-200
tools/run_bounded.sh 200 gpurun_out/bench${N}_dp.log $TR --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 --skip-e2e --cpu-steps 0 --parallel dp; echo "dp rc=$?"
grep '^{"metric' gpurun_out/bench${N}_dp.log | cut -c1-200
