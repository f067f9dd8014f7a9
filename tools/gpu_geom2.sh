#!/bin/bash
mkdir -p gpurun_out
QCS_TILE_SHRINK=2 timeout 900 python -m pytest tests/test_gpu_arms.py tests/test_gpu_parity.py -m gpu -q --no-header -x 2>&1 | tail -3
for cfg in "1 4" "2 4" "2 3"; do
  set -- $cfg
  QCS_TILE_SHRINK=$1 QCS_PIPE_CTAS11=$2 timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/bench_g2_t$1_c$2.json 2> gpurun_out/bench_g2_t$1_c$2.err
  QCS_TILE_SHRINK=$1 QCS_PIPE_CTAS11=$2 timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --dtype c128 --qubits 28 > gpurun_out/bench_g2c128_t$1_c$2.json 2> gpurun_out/bench_g2c128_t$1_c$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_g2*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'gates/s', round(d['ms_per_step'],1),'ms', d['config']['passes_per_step'],'passes', 'pergate', round(d['roofline_per_gate_pass']['ms_per_launch'],3), 'norm', d['norm_check'])
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-300:])
PY
