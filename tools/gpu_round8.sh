#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/bench_*
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
for mo in 6 16 32; do
  timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --max-ops $mo > gpurun_out/bench_mo$mo.json 2> gpurun_out/bench_mo$mo.err
done
timeout 600 python bench.py --steps 3 --warmup 3 --dtype c128 --qubits 28 --skip-cpu --skip-e2e > gpurun_out/bench_c128.json 2> gpurun_out/bench_c128.err
cat gpurun_out/summary.txt
tail -n 8 gpurun_out/pytest_gpu.log
