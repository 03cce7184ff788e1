#!/bin/bash
# round 2: the suite, both bench arms, the ncu launch list of the bench command and full captures of the tile kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|FAILED\|passed\|failed" gpurun_out/pytest_final.log | cut -c1-250 | head -12
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err | cut -c1-200
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_reference.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value %.5g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.5g" % d["e2e"]["value"], d["details"]["final_loss"], d["details"]["flash_fallbacks"], d["parity"], d["cpu_baseline"]["value"], d["clocks"])
print(d["roofline"]["all_ms"], d["roofline"]["frac"], d["roofline"]["step"])
x = d["extra"]["fb15k237"]; print("fb237", x["value"], x["ms_per_step"], x["e2e"]["value"], x["roofline"]["all_ms"])
PY
if [ "$1" == "ncu" ]; then
  timeout 300 python bench.py --steps 3 --warmup 3 --cpu-steps 0 --skip-extra --skip-e2e > gpurun_out/plain_launches.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
      python bench.py --steps 3 --warmup 3 --cpu-steps 0 --skip-extra --skip-e2e > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
  timeout 200 python tools/prof_tiles.py wd5m > gpurun_out/prof_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_bwd4 -s 2 -c 2 -o gpurun_out/prof_wd5m_final python tools/prof_tiles.py wd5m > gpurun_out/ncu_wd5m.log 2>&1; echo "ncu wd5m rc=$?"
  timeout 200 python tools/prof_tiles.py fb237 > gpurun_out/prof_plain_fb.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_bwd4 -s 2 -c 2 -o gpurun_out/prof_fb237_final python tools/prof_tiles.py fb237 > gpurun_out/ncu_fb237.log 2>&1; echo "ncu fb237 rc=$?"
fi
