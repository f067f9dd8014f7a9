#!/bin/bash
# one ncu --set full capture (with source) of a fused pass of the headline circuit
mkdir -p gpurun_out
export NCU_MAX_OPS=${NCU_MAX_OPS:-32}
timeout 300 python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/ncu_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 10 -c 2 -o /tmp/prof_fused python tools/ncu_target.py > gpurun_out/ncu_fused.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/prof_fused.ncu-rep --page source --csv > gpurun_out/prof_source_fused.csv 2>/dev/null
ncu -i /tmp/prof_fused.ncu-rep --page raw --csv > gpurun_out/prof_raw_fused.csv 2>/dev/null
ls -la gpurun_out/prof_*fused*
