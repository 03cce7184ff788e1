#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
tools/run_bounded.sh 240 gpurun_out/multi_gpu_check_$N.log $TR --master-port 29510 tests/multi_gpu_check.py; echo "multi_gpu_check rc=$?"
grep "==\|Error\|error\|assert\|unavailable" gpurun_out/multi_gpu_check_$N.log | head -20
tools/run_bounded.sh 200 gpurun_out/ns_sharded_check_$N.log $TR --master-port 29522 tests/ns_sharded_check.py; echo "ns_sharded_check rc=$?"
grep "==\|Error\|error\|assert" gpurun_out/ns_sharded_check_$N.log | head -12
tools/run_bounded.sh 200 gpurun_out/p2p_ns_check_$N.log $TR --master-port 29521 tests/p2p_ns_check.py; echo "p2p_ns_check rc=$?"
grep "==\|Error\|error\|assert" gpurun_out/p2p_ns_check_$N.log | head -12
