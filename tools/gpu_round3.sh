#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/*.ncu-rep gpurun_out/prof_* gpurun_out/bench_*
QCS_TEST_MODES=0 timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_pipe.log 2>&1
echo "pytest pipe exit $?" | tee -a gpurun_out/summary.txt
QCS_TEST_MODES=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -rf --timeout 300 -k "position or random_positions or high_and_low or lazy_norm or fused_pass_matches or circuit_run_fused" > gpurun_out/pytest_bulk.log 2>&1
echo "pytest bulk exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
timeout 900 python tools/microbench.py --n 30 --modes ${MB_MODES:-0} > gpurun_out/microbench.log 2>&1
echo "microbench exit $?" | tee -a gpurun_out/summary.txt
for mo in 6 10 16 24; do
  timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --max-ops $mo > gpurun_out/bench_mo$mo.json 2> gpurun_out/bench_mo$mo.err
done
timeout 900 python bench.py --steps 3 --warmup 3 --no-fuse --skip-cpu --skip-e2e > gpurun_out/bench_nofuse.json 2> gpurun_out/bench_nofuse.err
timeout 600 python bench.py --steps 3 --warmup 3 --dtype c128 --qubits 28 --skip-cpu --skip-e2e --max-ops 10 > gpurun_out/bench_c128.json 2> gpurun_out/bench_c128.err
timeout 300 python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -c 18 -o /tmp/prof python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?" | tee -a gpurun_out/summary.txt
ncu -i /tmp/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null
ncu -i /tmp/prof.ncu-rep --page source --csv --kernel-name regex:260 --launch-count 1 > gpurun_out/prof_source_fused.csv 2>gpurun_out/prof_source_fused.err
ncu -i /tmp/prof.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/prof_source_k1.csv 2>/dev/null
cat gpurun_out/summary.txt
for f in gpurun_out/pytest_pipe.log gpurun_out/pytest_bulk.log gpurun_out/smoke.log; do echo "== $f"; tail -n 8 $f; done
