#!/bin/bash
# final check of the round: full GPU suite, smoke, the bench lines that the docs quote
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 --dtype c128 --qubits 28 --skip-cpu --skip-e2e > gpurun_out/bench_c128.json 2> gpurun_out/bench_c128.err
timeout 900 python bench.py --workload qpe --steps 3 --warmup 3 > gpurun_out/bench_qpe32.json 2> gpurun_out/bench_qpe32.err
tail -n 3 gpurun_out/pytest_gpu.log
python - <<'PY'
import json,glob
for f in ('gpurun_out/bench_default.json','gpurun_out/bench_c128.json','gpurun_out/bench_qpe32.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), d['unit'], round(d['ms_per_step'],1),'ms', d['config'].get('passes_per_step'),'passes', 'frac', round(d['roofline']['frac'],3), 'pergate', round(d['roofline_per_gate_pass']['ms_per_launch'],3), 'e2e', (d.get('e2e') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
