#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_r10.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_r10.log | cut -c1-250 | head -12
timeout 600 python bench.py --steps 40 --warmup 3 --cpu-steps 0 --skip-extra > gpurun_out/bench_wd_r10.json 2> gpurun_out/bench_wd_r10.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_wd_r10.json").read().strip().splitlines()[-1])
print("1 GPU value %.4g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["details"]["final_loss"], {k.split(" ")[0]: round(x, 4) for k, x in d["roofline"]["all_ms"].items()})
PY
