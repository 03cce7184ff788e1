#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
tools/run_bounded.sh 200 gpurun_out/tl_p2p.log $TR --master-port 29513 bench.py --gpus 2 --parallel p2p --timeline; echo "rc=$?"
cp gpurun_out/timeline_bench.txt gpurun_out/timeline_p2p_rank0.txt; cp gpurun_out/timeline_bench_rank1.txt gpurun_out/timeline_p2p_rank1.txt
tools/run_bounded.sh 200 gpurun_out/tl_dp.log $TR --master-port 29514 bench.py --gpus 2 --parallel dp --timeline; echo "rc=$?"
cp gpurun_out/timeline_bench.txt gpurun_out/timeline_dp_rank0.txt
tail -22 gpurun_out/timeline_p2p_rank0.txt | cut -c1-140
