#!/bin/bash
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x -k "flash or large_table" 2>&1 | tail -2
tools/run_bounded.sh 240 gpurun_out/multi_gpu_check_$N.log $TR --master-port 29510 tests/multi_gpu_check.py; echo "multi_gpu_check rc=$?"
grep "row-sharded\|Error\|assert" gpurun_out/multi_gpu_check_$N.log | cut -c1-160 | head -8
tools/run_bounded.sh 300 gpurun_out/bench_wd5m_$N.log $TR --master-port 29513 bench.py --gpus $N --steps 30 --warmup 3; echo "bench rc=$?"
timeout 400 python bench.py --steps 30 --warmup 3 --cpu-steps 0 --skip-extra > gpurun_out/bench_wd_r13.json 2> gpurun_out/bench_wd_r13.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_wd5m_2.log", "gpurun_out/bench_wd_r13.json"):
    for line in open(f):
        if line.startswith('{"metric'):
            d = json.loads(line)
            print(f, "value %.5g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.5g" % d["e2e"]["value"], "loss", d["details"]["final_loss"], d["details"]["flash_fallbacks"], {k.split(" ")[0]: round(x, 4) for k, x in d["roofline"]["all_ms"].items()})
PY
