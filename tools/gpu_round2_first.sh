#!/bin/bash
# First GPU session of the next round: everything that was written after round 1's GPU budget was spent.
#   here (no GPU):   tools/build_variant.sh s128 tc_bwd -DKGEB_STR_ROWS=128
#                    tools/build_variant.sh gtmem tc_bwd -DKGEB_G_TMEM
#                    tools/build_variant.sh poly2 tc_bwd "-DKGEB_POLY8_BCE=2 -DKGEB_POLY8_STATS=2"
#   then:            gpurun --timeout 1500 -- bash tools/gpu_round2_first.sh
#   and (2 GPUs):    gpurun --gpus 2 --timeout 600 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
#                        --master-addr 127.0.0.1 --master-port 29521 tests/p2p_ns_check.py'
mkdir -p gpurun_out
# 1. the whole GPU suite; -rxX lists the staged (non-strict xfail) tests: XPASS -> remove the marker, XFAIL -> look
timeout 900 python -m pytest tests -m gpu -q -rxX > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
# 2. default bench, then the tuning builds: parity subset + device-resident value and kernel times
timeout 300 python bench.py --steps 30 --warmup 5 --cpu-steps 0 --skip-e2e > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
for v in s128 gtmem poly2; do
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
# 3. the Wikidata5M-shaped KL step with the tuning builds (tensor-side bound there)
for v in base s128 gtmem; do
  if [ $v == base ]; then L=$PWD/kge-1_b200/libkgeb200.so; else L=$PWD/kge-1_b200/variants/libkgeb200_$v.so; fi
  [ -f $L ] || continue
  KGEB200_LIB=$L timeout 300 python bench_extra.py --workload wd5m-1vsall --batch 1024 --graph-step > gpurun_out/extra_1vsall_$v.log 2>&1
  echo "$v:"; grep '^{"workload' gpurun_out/extra_1vsall_$v.log | cut -c1-200
done
