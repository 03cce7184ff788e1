#!/bin/bash
mkdir -p gpurun_out
echo "== parity (A in TMEM)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_aux.py -m gpu -q > gpurun_out/pytest_r7.log 2>&1; echo "rc=$?"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_r7.log | cut -c1-250 | head -12
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench_fb237.py --steps 30 --warmup 5 --cpu-steps 0 --skip-e2e > gpurun_out/bench_fb_$name.json 2> gpurun_out/bench_fb_$name.err
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --cpu-steps 0 --skip-extra --skip-e2e > gpurun_out/bench_wd_$name.json 2> gpurun_out/bench_wd_$name.err
  python - <<PY
import json
for f in ("gpurun_out/bench_fb_$name.json", "gpurun_out/bench_wd_$name.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("$name", f.split("_")[1], "value %.4g" % d["value"], "ms %.4f" % d["ms_per_step"], {k.split(" ")[0]: round(x, 4) for k, x in d["roofline"]["all_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-600:])
PY
}
run atmem A=1
run noatmem KGEB_NO_A_TMEM=1
run poly2 KGEB200_LIB=$PWD/kge-1_b200/variants/libkgeb200_poly2.so
run poly3 KGEB200_LIB=$PWD/kge-1_b200/variants/libkgeb200_poly3.so
