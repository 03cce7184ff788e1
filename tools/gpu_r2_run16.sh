#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_aux.py tests/test_gpu_parity.py -m gpu -q -x -k "negative or segment or fused_negative" > gpurun_out/pytest_ns.log 2>&1; echo "rc=$?"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_ns.log | cut -c1-250 | head
for v in "--no-fused-slot" "" "--atomic"; do
  timeout 200 python bench_extra.py --workload wnrr-rotate-ns --batch 512 --graph-step --steps 100 $v --kernels > gpurun_out/extra_ns$v.log 2>&1; grep '^{"workload' "gpurun_out/extra_ns$v.log" | cut -c1-420
done
head -14 gpurun_out/kernels_wnrr_rotate_ns_graph.txt | cut -c1-150; head -12 gpurun_out/kernels_wnrr_rotate_ns_graph_atomic.txt | cut -c1-150
for m in complex transe; do
  timeout 900 python bench_extra.py --workload wd5m-eval-$m --batch 512 --real-filter > gpurun_out/extra_eval_$m.log 2>&1; grep '^{"workload' gpurun_out/extra_eval_$m.log | cut -c1-700; tail -2 gpurun_out/extra_eval_$m.log | cut -c1-200
done
