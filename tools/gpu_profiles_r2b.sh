#!/bin/bash
# round-2 evidence for the FINAL tensor-core kernel (fp16 operands, four groups): launch list of the bench command, one
# --set full capture (with source) on passes of the headline circuit, per-pass times
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-strong --skip-configs"
timeout 600 $BENCH > gpurun_out/r2_bench_for_launches.json 2> gpurun_out/r2_bench_for_launches.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_ncu.csv $BENCH > gpurun_out/r2_launches_bench.log 2>&1
echo "launch list exit $?"
timeout 300 python tools/tc_ncu_real_target.py > gpurun_out/r2_ncu_tc_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_kernel_tc -s 1 -c 1 -o /tmp/prof_tc_real python tools/tc_ncu_real_target.py > gpurun_out/r2_ncu_tc.log 2>&1
echo "ncu full exit $?"
cat gpurun_out/r2_ncu_tc_plain.log
ncu -i /tmp/prof_tc_real.ncu-rep --page source --csv > gpurun_out/r2_ncu_tc_source.csv 2>/dev/null
ncu -i /tmp/prof_tc_real.ncu-rep --page raw --csv > gpurun_out/r2_ncu_tc_raw.csv 2>/dev/null
timeout 300 python tools/pass_times.py > gpurun_out/r2_pass_times.log 2>&1; cp gpurun_out/pass_times.json gpurun_out/r2_pass_times.json
tail -3 gpurun_out/r2_pass_times.log
ls -la gpurun_out/r2_*
