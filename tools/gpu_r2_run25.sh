#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -rf > gpurun_out/pytest_all.log 2>&1; rc=$?; echo "pytest rc=$rc"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_all.log | cut -c1-300 | head -12
[ $rc -ne 0 ] && exit 1
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value %.5g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.5g" % d["e2e"]["value"], d["details"]["final_loss"], d["details"]["flash_fallbacks"], d.get("parity"), d["cpu_baseline"]["value"], d["clocks"])
print(d["roofline"]["kernel"][:60], d["roofline"]["all_ms"], d["roofline"]["frac"], d["roofline"]["step"])
x = d["extra"]["fb15k237"]; print("fb237", x["value"], x["ms_per_step"], x["e2e"]["value"], x["roofline"]["all_ms"])
PY
