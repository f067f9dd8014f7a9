#!/bin/bash
# First GPU session: parity (plain staging first, TMA variant separately so a hang cannot hide the rest),
# smoke, per-bit microbenchmark, first bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
QCS_TEST_MODES=0 timeout 1200 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_plain.log 2>&1
echo "pytest plain exit $?" | tee -a gpurun_out/summary.txt
QCS_TEST_MODES=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -rf --timeout 300 -k "position or random_positions or high_and_low or lazy_norm or fused_pass_matches or circuit_run_fused" > gpurun_out/pytest_bulk.log 2>&1
echo "pytest bulk exit $?" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
timeout 900 python tools/microbench.py --n 30 > gpurun_out/microbench.log 2>&1
echo "microbench exit $?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/summary.txt
QCS_TILE_MODE=bulk timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/bench_bulk.json 2> gpurun_out/bench_bulk.err
echo "bench bulk exit $?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest_plain.log gpurun_out/pytest_bulk.log gpurun_out/smoke.log
tail -3 gpurun_out/bench.json gpurun_out/bench_bulk.json
