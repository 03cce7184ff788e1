#!/bin/bash
# Adagrad in the dTable flush: new tests, the stepper tests, NS slot kernel check, WD5M bench with / without the fused update
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x -k "touched or adagrad_in or large_table" > gpurun_out/pytest_upd.log 2>&1; echo "upd rc=$?"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_upd.log | cut -c1-300 | head -20
timeout 600 python -m pytest tests/test_gpu_aux.py -m gpu -q -x -k "fused_negative or negative_sampling" > gpurun_out/pytest_ns.log 2>&1; echo "ns rc=$?"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_ns.log | cut -c1-250 | head
KGEB_NO_FUSED_UPDATE=1 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_wd_noupd.json 2> gpurun_out/bench_wd_noupd.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_wd_noupd.json
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_wd_upd.json 2> gpurun_out/bench_wd_upd.err; echo "rc=$?"; cut -c1-2500 gpurun_out/bench_wd_upd.json; tail -3 gpurun_out/bench_wd_upd.err
for v in "" "--atomic"; do
  timeout 200 python bench_extra.py --workload wnrr-rotate-ns --batch 512 --graph-step --steps 100 $v --kernels > gpurun_out/extra_ns2$v.log 2>&1; grep '^{"workload' "gpurun_out/extra_ns2$v.log" | cut -c1-420
done
head -5 gpurun_out/kernels_wnrr_rotate_ns_graph_atomic.txt | cut -c1-150
