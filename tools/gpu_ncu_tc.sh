#!/bin/bash
# one ncu --set full capture (with source) of the tensor-core tile kernel on a synthetic four-block pass
mkdir -p gpurun_out
export NCU_VARIANTS=${NCU_VARIANTS:-low}
timeout 300 python tools/tc_ncu_target.py > gpurun_out/ncu_tc_plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/ncu_tc_plain.log; exit 1; }
cat gpurun_out/ncu_tc_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_kernel_tc -s 1 -c 1 -o /tmp/prof_tc python tools/tc_ncu_target.py > gpurun_out/ncu_tc.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/prof_tc.ncu-rep --page source --csv > gpurun_out/prof_source_tc.csv 2>/dev/null
ncu -i /tmp/prof_tc.ncu-rep --page raw --csv > gpurun_out/prof_raw_tc.csv 2>/dev/null
ls -la gpurun_out/prof_*tc.csv
