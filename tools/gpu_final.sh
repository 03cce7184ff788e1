#!/bin/bash
# usage (from the repo root): gpurun --timeout 2400 -- bash tools/gpu_final.sh
# Round-end style validation on one B200: parity tests, smoke, both bench arms, ncu launch list + full capture, other workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 200 --csv --log-file gpurun_out/launches_r1s2_bench_default.csv \
  python bench.py --steps 5 --warmup 3 --skip-e2e --cpu-steps 0 > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_bwd_kernel -s 2 -c 2 -o gpurun_out/prof_bwd_final \
  python bench.py --no-graph --steps 1 --warmup 3 --skip-e2e --cpu-steps 0 > gpurun_out/ncu_bwd_final.log 2>&1; echo "ncu full rc=$?"
rm -f gpurun_out/extra_workloads.jsonl
timeout 400 python bench_extra.py --workload wd5m-1vsall --batch 1024 --graph-step > gpurun_out/extra_1vsall.log 2>&1
timeout 300 python bench_extra.py --workload wd5m-1vsall --batch 256 --graph-step > gpurun_out/extra_1vsall256.log 2>&1
timeout 300 python bench_extra.py --workload wd5m-eval-complex --batch 512 --math tf32 > gpurun_out/extra_evalc.log 2>&1
timeout 300 python bench_extra.py --workload wd5m-eval-transe --batch 128 > gpurun_out/extra_evalt.log 2>&1
timeout 300 python bench_extra.py --workload wnrr-rotate-ns --batch 512 --graph-step > gpurun_out/extra_ns.log 2>&1
cut -c1-260 gpurun_out/extra_workloads.jsonl
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('value %.2fM'%(d['value']/1e6), 'ms %.4f'%d['ms_per_step'], 'e2e %.2fM'%(d['e2e']['value']/1e6), 'dc %.2fM'%(d['e2e_device_collate']['value']/1e6), d['roofline']['all_ms'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])
PY
