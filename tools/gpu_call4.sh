#!/bin/bash
# 2-GPU session: peer-memory data-parallel step against the NCCL one; 2-GPU bench in both modes
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
tools/run_bounded.sh 300 gpurun_out/p2p_check.log $TR --master-port 29511 tests/p2p_check.py; echo "p2p_check rc=$?"
grep -v "^W\|^\[W\|Warning" gpurun_out/p2p_check.log | tail -12
tools/run_bounded.sh 240 gpurun_out/bench2_dp.log $TR --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --skip-e2e --cpu-steps 0 --parallel dp; echo "dp rc=$?"
grep '^{"metric' gpurun_out/bench2_dp.log | cut -c1-220
tools/run_bounded.sh 240 gpurun_out/bench2_p2p.log $TR --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --skip-e2e --cpu-steps 0 --parallel p2p; echo "p2p rc=$?"
grep '^{"metric' gpurun_out/bench2_p2p.log | cut -c1-220; tail -5 gpurun_out/bench2_p2p.log | cut -c1-300
