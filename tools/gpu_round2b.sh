#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/*.ncu-rep
QCS_TEST_MODES=0 timeout 600 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 -k "golden_measurements" > gpurun_out/pytest_pipe.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/summary.txt
timeout 900 python tools/microbench.py --n 30 > gpurun_out/microbench.log 2>&1
echo "microbench exit $?" | tee -a gpurun_out/summary.txt
for mo in 6 10 16 24; do
  timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --max-ops $mo > gpurun_out/bench_mo$mo.json 2> gpurun_out/bench_mo$mo.err
done
QCS_PIPE_CTAS=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --max-ops 10 > gpurun_out/bench_ctas1.json 2> gpurun_out/bench_ctas1.err
QCS_PIPE_STAGES=2 timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --max-ops 10 > gpurun_out/bench_stages2.json 2> gpurun_out/bench_stages2.err
timeout 900 python bench.py --steps 3 --warmup 3 --no-fuse --skip-cpu --skip-e2e > gpurun_out/bench_nofuse.json 2> gpurun_out/bench_nofuse.err
timeout 300 python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -c 18 -o /tmp/prof_r1 python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?" | tee -a gpurun_out/summary.txt
ls -la /tmp/prof_r1.ncu-rep >> gpurun_out/summary.txt
ncu -i /tmp/prof_r1.ncu-rep --page raw --csv > gpurun_out/prof_r1_raw.csv 2>/dev/null
ncu -i /tmp/prof_r1.ncu-rep --page details --csv > gpurun_out/prof_r1_details.csv 2>/dev/null
ncu -i /tmp/prof_r1.ncu-rep --page source --csv --kernel-id :::10 > gpurun_out/prof_r1_source_fused.csv 2>/dev/null
ncu -i /tmp/prof_r1.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/prof_r1_source_single.csv 2>/dev/null
du -sh gpurun_out >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 4 gpurun_out/pytest_pipe.log
