#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf -x > gpurun_out/pytest_run3.log 2>&1; echo "pytest rc=$?"
grep -v "^\[\|^  \|Warning\|^$" gpurun_out/pytest_run3.log | grep -n "^E \|FAILED\|passed\|failed\|Error" | head -40
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_wd5m.json 2> gpurun_out/bench_wd5m.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_wd5m.err; cut -c1-2500 gpurun_out/bench_wd5m.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_wd5m_ref.json 2> gpurun_out/bench_wd5m_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/bench_wd5m_ref.err; cut -c1-1200 gpurun_out/bench_wd5m_ref.json
free -g | head -2; nproc
