#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/bench_* gpurun_out/prof_* gpurun_out/launches*
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
for mo in 6 16 24 32 48; do
  timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --max-ops $mo > gpurun_out/bench_mo$mo.json 2> gpurun_out/bench_mo$mo.err
done
timeout 600 python bench.py --steps 3 --warmup 3 --dtype c128 --qubits 28 --skip-cpu --skip-e2e > gpurun_out/bench_c128.json 2> gpurun_out/bench_c128.err
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 && \
NCU_MAX_OPS=${NCU_MAX_OPS:-16} timeout 900 ncu --set full --clock-control none -k regex:tile_kernel -c 18 -o /tmp/prof python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?" | tee -a gpurun_out/summary.txt
ncu -i /tmp/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null
NCU_MAX_OPS=${NCU_MAX_OPS:-16} timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 10 -c 1 -o /tmp/prof_fused python tools/ncu_target.py > gpurun_out/ncu_fused.log 2>&1
ncu -i /tmp/prof_fused.ncu-rep --page source --csv > gpurun_out/prof_source_fused.csv 2>/dev/null
cat gpurun_out/summary.txt
tail -n 6 gpurun_out/pytest_gpu.log
