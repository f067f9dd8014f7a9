#!/bin/bash
# v2 register ops: parity, then bench at several fusion depths
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/bench_* gpurun_out/prof_* gpurun_out/launches*
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
for mo in 8 16 32 48 64; do
  timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --max-ops $mo > gpurun_out/bench_mo$mo.json 2> gpurun_out/bench_mo$mo.err
done
timeout 600 python bench.py --steps 3 --warmup 3 --dtype c128 --qubits 28 --skip-cpu --skip-e2e --max-ops 32 > gpurun_out/bench_c128.json 2> gpurun_out/bench_c128.err
cat gpurun_out/summary.txt
tail -n 6 gpurun_out/pytest_gpu.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'gates/s', round(d['ms_per_step'],1),'ms', d['config']['passes_per_step'],'passes', round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
