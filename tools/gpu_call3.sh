#!/bin/bash
# GPU session: parity after the issuer rewrite; bench; the WD5M workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_elect.json 2> gpurun_out/bench_elect.err; tail -c 1500 gpurun_out/bench_elect.json
rm -f gpurun_out/extra_workloads.jsonl
timeout 400 python bench_extra.py --workload wd5m-1vsall --batch 1024 --graph-step > gpurun_out/extra_1vsall.log 2>&1; tail -2 gpurun_out/extra_1vsall.log
timeout 300 python bench_extra.py --workload wd5m-eval-complex --batch 512 --math tf32 > gpurun_out/extra_evalc.log 2>&1; tail -2 gpurun_out/extra_evalc.log
