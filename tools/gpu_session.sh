#!/bin/bash
# round-end style session: parity, smoke, bench lines (headline, config 2, config 3, reference arm), ncu evidence
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/bench_* gpurun_out/prof_* gpurun_out/launches*
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 --dtype c128 --qubits 28 --skip-cpu --skip-e2e > gpurun_out/bench_c128.json 2> gpurun_out/bench_c128.err
timeout 900 python bench.py --workload qpe --steps 3 --warmup 3 > gpurun_out/bench_qpe32.json 2> gpurun_out/bench_qpe32.err
echo "bench qpe exit $?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" | tee -a gpurun_out/summary.txt
export NCU_MAX_OPS=128
timeout 300 python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:tile_kernel -c 18 -o /tmp/prof python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?" | tee -a gpurun_out/summary.txt
ncu -i /tmp/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 10 -c 1 -o /tmp/prof_fused python tools/ncu_target.py > gpurun_out/ncu_fused.log 2>&1
ncu -i /tmp/prof_fused.ncu-rep --page source --csv > gpurun_out/prof_source_fused.csv 2>/dev/null
timeout 300 python tools/pass_times.py > gpurun_out/pass_times.log 2>&1
timeout 300 python tools/microbench_misc.py > gpurun_out/microbench_misc.log 2>&1
cat gpurun_out/summary.txt
tail -n 4 gpurun_out/pytest_gpu.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],3), d['unit'], round(d['ms_per_step'],1),'ms', d['config'].get('passes_per_step'),'passes', (d.get('roofline') or {}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
