#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x -k "touched or adagrad_in or large_table" > gpurun_out/pytest_upd.log 2>&1; rc=$?; echo "upd rc=$rc"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_upd.log | cut -c1-300 | head -20

for v in 1 0; do
KGEB_NO_FUSED_UPDATE=$v timeout 300 python bench.py --steps 20 --warmup 3 --cpu-steps 0 --skip-extra > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ab_$v.json").read().strip().splitlines()[-1])
print("no_fused_update=$v", "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d["details"]["final_loss"], d["clocks"])
PY
done
