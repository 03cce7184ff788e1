#!/bin/bash
mkdir -p gpurun_out
for v in base rg4 rg4lg8 rg2lg8 rg4lg16; do
  if [ $v == base ]; then L=$PWD/kge-1_b200/libkgeb200.so; else L=$PWD/kge-1_b200/variants/libkgeb200_$v.so; fi
  KGEB200_LIB=$L timeout 200 python bench.py --steps 30 --warmup 5 --cpu-steps 0 --skip-e2e > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$v.json').read().strip().splitlines()[-1])
    print('$v', 'value %.2fM'%(d['value']/1e6), 'ms %.4f'%d['ms_per_step'], {k.split(' ')[0]:round(x*1e3,1) for k,x in d['roofline']['all_ms'].items()}, 'loss', d['config']['final_loss'])
except Exception as e:
    print('$v ERR', e)
PY
done
KGEB200_LIB=$PWD/kge-1_b200/variants/libkgeb200_rg4lg8.so timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
