#!/bin/bash
# last sanity of round 2: the shape / drop-in tests and a short bench line (e2e through the reference's job object)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_shapes.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/pytest_last.log 2>&1; rc=$?; echo "rc=$rc"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_last.log | cut -c1-400 | head -20
[ $rc -ne 0 ] && exit 1
timeout 300 python bench.py --skip-extra --cpu-steps 0 > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_last.json").read().strip().splitlines()[-1])
print("value %.5g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.5g" % d["e2e"]["value"], d["details"]["final_loss"], d["clocks"], d["gpu_launches_per_step"])
PY
