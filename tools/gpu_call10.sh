#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
tools/run_bounded.sh 400 gpurun_out/multi_gpu_check.log $TR --master-port 29515 tests/multi_gpu_check.py; echo "multi_gpu_check rc=$?"
grep "==\|Error\|assert" gpurun_out/multi_gpu_check.log | head -12
CUDA_VISIBLE_DEVICES=0 timeout 400 ncu --kernel-name-base demangled -k regex:"kgeb|label_weight|loss_rows" --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
  --log-file gpurun_out/launches_r1s2_bench_default.csv python bench.py --steps 5 --warmup 3 --skip-e2e --cpu-steps 0 > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
grep -c "tc_bwd" gpurun_out/launches_r1s2_bench_default.csv
