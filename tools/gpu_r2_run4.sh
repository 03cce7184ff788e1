#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_run4.log 2>&1; echo "pytest rc=$?"
grep -v "^\[\|^  \|Warning\|^$" gpurun_out/pytest_run4.log | grep -n "^E  \|FAILED\|passed\|failed\|Error" | cut -c1-300 | head -40
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-steps 0 --skip-extra > gpurun_out/bench_wd5m_flash.json 2> gpurun_out/bench_wd5m_flash.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_wd5m_flash.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_wd5m_flash.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'loss',d['details']['final_loss'],'launches',d['gpu_launches_per_step'])
print(d['roofline']['all_ms'], 'frac', d['roofline']['frac'], 'step', d['roofline']['step']['frac'], d['clocks'])
PY
timeout 300 python bench_extra.py --workload wd5m-1vsall --batch 1024 --graph-step --kernels > gpurun_out/extra_1vsall_flash.log 2>&1; grep '^{"workload' gpurun_out/extra_1vsall_flash.log | cut -c1-200; head -12 gpurun_out/kernels_wd5m_1vsall_B1024.txt | cut -c1-200
