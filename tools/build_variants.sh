#!/bin/bash
# A/B builds of the split kernels: one library per combination of the NRP_V_* switches (partition.cuh), written to
# _variants/ (git-ignored; travels to the GPU box).  Select one at run time with NRPCALC_B200_LIB=<path> (nrpcalc_b200/_native.py).
#   tools/build_variants.sh "name:DUMP,LATE,EARLY,FPIPE" ...
set -eu
cd "$(dirname "$0")/../nrpcalc_b200/csrc"
OUT=../../_variants
mkdir -p $OUT
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wno-unused-function --expt-relaxed-constexpr"
for spec in "$@"; do
  name=${spec%%:*}; IFS=, read -r d l y f <<< "${spec#*:}"
  ( nvcc $FLAGS -DNRP_V_DUMP=$d -DNRP_V_LATE=$l -DNRP_V_EARLY=$y -DNRP_V_FPIPE=${f:-1} -shared -o $OUT/libnrpb200_$name.so nrp_b200.cu -lcudart && echo "built $name" ) &
done
wait
