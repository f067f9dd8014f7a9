#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/*.ncu-rep gpurun_out/bench_*
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" | tee -a gpurun_out/summary.txt
QCS_PIPE_CTAS=3 QCS_TEST_MODES=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_ctas3.log 2>&1
echo "pytest ctas3 exit $?" | tee -a gpurun_out/summary.txt
for c in 2 3; do for mo in 6 16 32; do
  QCS_PIPE_CTAS=$c timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --max-ops $mo > gpurun_out/bench_c${c}_mo$mo.json 2> gpurun_out/bench_c${c}_mo$mo.err
done; done
QCS_PIPE_CTAS=3 timeout 600 python bench.py --steps 3 --warmup 3 --dtype c128 --qubits 28 --skip-cpu --skip-e2e > gpurun_out/bench_c3_c128.json 2> gpurun_out/bench_c3_c128.err
timeout 300 python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 && \
NCU_MAX_OPS=16 timeout 900 ncu --set full --clock-control none -k regex:tile_kernel -c 18 -o /tmp/prof python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?" | tee -a gpurun_out/summary.txt
ncu -i /tmp/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null
cat gpurun_out/summary.txt
tail -n 8 gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/pytest_ctas3.log
