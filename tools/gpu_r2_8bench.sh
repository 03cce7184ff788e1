#!/bin/bash
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
tools/run_bounded.sh 200 gpurun_out/bench_wd5m_$N.log $TR --master-port 29540 bench.py --gpus $N --steps 30 --warmup 3; echo "bench $N rc=$?"
grep '^{"metric' gpurun_out/bench_wd5m_$N.log | cut -c1-330; grep -i "error\|unavailable" gpurun_out/bench_wd5m_$N.log | head -3
