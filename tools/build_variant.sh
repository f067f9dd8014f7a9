#!/bin/bash
# Scratch build of libqcs with extra nvcc flags (timing experiments): tools/build_variant.sh <name> <flags...>
# -> build/<name>/libqcs.so (git-ignored, travels to the GPU box); select it with QCS_LIB=build/<name>/libqcs.so
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/build/$name
mkdir -p $out/obj
cd $root/qcircuits_b200/csrc
[ -f qcs_regops.inc ] || python3 gen_regops.py > qcs_regops.inc
for f in qcs_api qcs_tile qcs_misc qcs_dist; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c $f.cu -o $out/obj/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libqcs.so $out/obj/*.o -lcudart
echo built $out/libqcs.so
