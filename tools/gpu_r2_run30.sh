#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v "^nvcc\|^built" | tail -6; echo "smoke rc=${PIPESTATUS[0]}"
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value %.5g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.5g" % d["e2e"]["value"], d["details"]["final_loss"], d["clocks"])
print(d["extra"].get("reference_cuda"))
PY
timeout 200 python bench_extra.py --workload wnrr-rotate-ns --batch 512 --graph-step --steps 20 --atomic > gpurun_out/ns_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:ns_fused_kernel -s 4 -c 1 -o gpurun_out/prof_ns_fused python bench_extra.py --workload wnrr-rotate-ns --batch 512 --graph-step --steps 20 --atomic > gpurun_out/ncu_ns.log 2>&1; echo "ncu ns rc=$?"; tail -2 gpurun_out/ncu_ns.log | cut -c1-200
