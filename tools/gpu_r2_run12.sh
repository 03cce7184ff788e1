#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 40 --warmup 3 --cpu-steps 0 --skip-extra > gpurun_out/bench_wd_r12.json 2> gpurun_out/bench_wd_r12.err; echo "rc=$?"; tail -3 gpurun_out/bench_wd_r12.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_wd_r12.json").read().strip().splitlines()[-1])
print("1 GPU value %.4g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e", d["e2e"]["value"], d["details"], {k.split(" ")[0]: round(x, 4) for k, x in d["roofline"]["all_ms"].items()}, d["roofline"]["frac"], d["roofline"]["step"]["frac"])
PY
