#!/bin/bash
# flexible low/high tile split: parity, then bench at several fusion depths and low-run shrinks
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/bench_*
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" | tee -a gpurun_out/summary.txt
for ls in 0 2 3; do
for mo in 32 48 64; do
  QCS_LOW_SHRINK=$ls timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --max-ops $mo > gpurun_out/bench_ls${ls}_mo$mo.json 2> gpurun_out/bench_ls${ls}_mo$mo.err
done
done
timeout 600 python bench.py --steps 3 --warmup 3 --dtype c128 --qubits 28 --skip-cpu --skip-e2e --max-ops 48 > gpurun_out/bench_c128.json 2> gpurun_out/bench_c128.err
PT_MAX_OPS=64 timeout 300 python tools/pass_times.py
tail -n 4 gpurun_out/pytest_gpu.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'gates/s', round(d['ms_per_step'],1),'ms', d['config']['passes_per_step'],'passes', round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
