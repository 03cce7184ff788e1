#!/bin/bash
mkdir -p gpurun_out
echo "== parity (colk smem)"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -q -x > gpurun_out/pytest_r8.log 2>&1; echo "rc=$?"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_r8.log | cut -c1-250 | head -8
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-steps 0 --skip-extra --skip-e2e > gpurun_out/bench_wd_colk.json 2> gpurun_out/bench_wd_colk.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_wd_colk.json").read().strip().splitlines()[-1])
print("1 GPU value %.4g" % d["value"], "ms %.4f" % d["ms_per_step"], {k.split(" ")[0]: round(x, 4) for k, x in d["roofline"]["all_ms"].items()})
PY
bash tools/gpu_multi_r2.sh 2
