#!/bin/bash
# quick check of a kernel change: parity of every interpreter arm + the circuit-level tests, then the headline bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_arms.py tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q --no-header -x 2>&1 | tail -4
for mo in 64; do
  timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --max-ops $mo > gpurun_out/bench_quick_mo$mo.json 2> gpurun_out/bench_quick_mo$mo.err
done
timeout 600 python bench.py --steps 3 --warmup 3 --dtype c128 --qubits 28 --skip-cpu --skip-e2e > gpurun_out/bench_quick_c128.json 2> gpurun_out/bench_quick_c128.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_quick_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'gates/s', round(d['ms_per_step'],1),'ms', d['config']['passes_per_step'],'passes', 'per-gate pass ms', round(d['roofline_per_gate_pass']['ms_per_launch'],3))
    except Exception as e: print(f,'ERR',e)
PY
