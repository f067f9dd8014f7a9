#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_arms.py tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_reference_properties.py -m gpu -q --no-header -x 2>&1 | tail -4
timeout 600 python bench.py --workload qpe --steps 3 --warmup 3 > gpurun_out/bench_qpe32.json 2> gpurun_out/bench_qpe32.err
timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_qpe32.json','gpurun_out/bench_quick.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), 'gates/s', round(d['ms_per_step'],1),'ms', d['config']['passes_per_step'],'passes', d.get('qpe_check'), d['norm_check'])
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-400:])
PY
