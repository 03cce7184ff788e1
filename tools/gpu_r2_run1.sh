#!/bin/bash
# round 2, GPU call 1: unmasked GPU suite, baseline bench + tuning builds, WD5M step kernel table, NS segment backward
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rxXf > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 30 --warmup 5 --cpu-steps 0 --skip-e2e > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
for v in s128 gtmem; do
  L=$PWD/kge-1_b200/variants/libkgeb200_$v.so
  [ -f $L ] || continue
  KGEB200_LIB=$L timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused or tensor_tiles or graph or stepper" > gpurun_out/pytest_$v.log 2>&1
  echo "$v parity rc=$?"; tail -2 gpurun_out/pytest_$v.log
  KGEB200_LIB=$L timeout 200 python bench.py --steps 30 --warmup 5 --cpu-steps 0 --skip-e2e > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1][:-5].ljust(8), "value %.2fM" % (d["value"] / 1e6), "ms %.4f" % d["ms_per_step"],
              {k.split(" ")[0]: round(x * 1e3, 1) for k, x in d["roofline"]["all_ms"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
for v in base s128 gtmem; do
  if [ $v == base ]; then L=$PWD/kge-1_b200/libkgeb200.so; K=--kernels; else L=$PWD/kge-1_b200/variants/libkgeb200_$v.so; K=; fi
  [ -f $L ] || continue
  KGEB200_LIB=$L timeout 300 python bench_extra.py --workload wd5m-1vsall --batch 1024 --graph-step $K > gpurun_out/extra_1vsall_$v.log 2>&1
  echo "$v:"; grep '^{"workload' gpurun_out/extra_1vsall_$v.log | cut -c1-260
done
head -30 gpurun_out/kernels_wd5m_1vsall_B1024.txt
timeout 200 python bench_extra.py --workload wnrr-rotate-ns --batch 512 --graph-step --kernels --steps 50 > gpurun_out/extra_ns.log 2>&1; grep '^{"workload' gpurun_out/extra_ns.log | cut -c1-300
timeout 200 python bench_extra.py --workload wnrr-rotate-ns --batch 512 --graph-step --segment-bwd --kernels --steps 50 > gpurun_out/extra_ns_seg.log 2>&1; grep '^{"workload' gpurun_out/extra_ns_seg.log | cut -c1-300
head -30 gpurun_out/kernels_wnrr_rotate_ns_graph.txt; head -30 gpurun_out/kernels_wnrr_rotate_ns_graph_seg.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
