#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 5" "1 6"; do
  set -- $cfg
  QCS_TILE_SHRINK=$1 QCS_LOW_SHRINK=$2 timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/bench_geom_t$1_s$2.json 2> gpurun_out/bench_geom_t$1_s$2.err
  QCS_TILE_SHRINK=$1 QCS_LOW_SHRINK=$2 timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --dtype c128 --qubits 28 > gpurun_out/bench_geom128_t$1_s$2.json 2> gpurun_out/bench_geom128_t$1_s$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_geom*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'gates/s', round(d['ms_per_step'],1),'ms', d['config']['passes_per_step'],'passes', 'norm', d['norm_check'])
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-300:])
PY
