#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_arms.py tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q --no-header -x 2>&1 | tail -3
for mo in 64 96 128; do
  timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --max-ops $mo > gpurun_out/bench_mo$mo.json 2> gpurun_out/bench_mo$mo.err
  timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --max-ops $mo --dtype c128 --qubits 28 > gpurun_out/bench128_mo$mo.json 2> gpurun_out/bench128_mo$mo.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench*_mo*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'gates/s', round(d['ms_per_step'],1),'ms', d['config']['passes_per_step'],'passes', 'norm', d['norm_check'])
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-300:])
PY
