#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_split.json').read().strip().splitlines()[-1])
print('value %.2fM'%(d['value']/1e6), 'ms %.4f'%d['ms_per_step'], 'e2e %.2fM'%(d['e2e']['value']/1e6), 'dc %.2fM'%(d['e2e_device_collate']['value']/1e6), d['roofline']['all_ms'], d['cpu_baseline']['value'])
PY
python bench.py --timeline > gpurun_out/tl1.log 2>&1; cut -c1-120 gpurun_out/timeline_bench.txt | grep -n "tc_bwd\|adagrad\|label_entry\|phase2" 
