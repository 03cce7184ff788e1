#!/bin/bash
mkdir -p gpurun_out
timeout 170 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value %.5g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.5g" % d["e2e"]["value"], d["details"]["final_loss"], d["clocks"], d["gpu_launches_per_step"], d["parity"]["rel_diff"], d["cpu_baseline"]["value"], d["extra"]["reference_cuda"]["ms_per_step"], d["extra"]["fb15k237"]["value"])
PY
