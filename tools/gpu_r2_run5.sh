#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/debug_flash.py 2>&1 | tail -12
echo "== v2 parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_aux.py -m gpu -q -x -k "not flash" > gpurun_out/pytest_v2.log 2>&1; echo "rc=$?"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_v2.log | cut -c1-250 | head -12
for v in 1 0; do
  echo "== KGEB_BWD_V2=$v"
  KGEB_BWD_V2=$v timeout 300 python bench_fb237.py --steps 30 --warmup 5 --cpu-steps 0 --skip-e2e > gpurun_out/bench_fb_v$v.json 2> gpurun_out/bench_fb_v$v.err
  KGEB_BWD_V2=$v timeout 600 python bench.py --steps 10 --warmup 3 --cpu-steps 0 --skip-extra --skip-e2e > gpurun_out/bench_wd_v$v.json 2> gpurun_out/bench_wd_v$v.err
  python - <<PY
import json
for f in ("gpurun_out/bench_fb_v$v.json", "gpurun_out/bench_wd_v$v.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "ms %.4f" % d["ms_per_step"], {k.split(" ")[0]: round(x, 4) for k, x in d["roofline"]["all_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-600:])
PY
done
