#!/bin/bash
# one ncu --set full capture (with source) of the tensor-core tile kernel on a real pass of the headline circuit
mkdir -p gpurun_out
timeout 300 python tools/tc_ncu_real_target.py > gpurun_out/ncu_tc_real_plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/ncu_tc_real_plain.log; exit 1; }
cat gpurun_out/ncu_tc_real_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_kernel_tc -s 1 -c 1 -o /tmp/prof_tc python tools/tc_ncu_real_target.py > gpurun_out/ncu_tc_real.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/prof_tc.ncu-rep --page source --csv > gpurun_out/prof_source_tc_real.csv 2>/dev/null
ncu -i /tmp/prof_tc.ncu-rep --page raw --csv > gpurun_out/prof_raw_tc_real.csv 2>/dev/null
ls -la gpurun_out/prof_*tc_real.csv
