#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -rf > gpurun_out/pytest_final.log 2>&1; rc=$?; echo "pytest rc=$rc"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_final.log | cut -c1-300 | head -12
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value %.5g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.5g" % d["e2e"]["value"], d["details"]["final_loss"], d["clocks"], d["gpu_launches_per_step"])
print(d["roofline"]["all_ms"], d["roofline"]["frac"], d["roofline"]["step"]["frac"], d["extra"]["reference_cuda"]["ms_per_step"], d["extra"]["fb15k237"]["value"])
PY
