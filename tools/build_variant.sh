#!/bin/bash
# One A/B build of the CUDA library with extra compile-time switches, written to _variants/ (git-ignored; travels to the GPU box).
# Select it at run time with NRPCALC_B200_LIB=_variants/libnrpb200_<name>.so (nrpcalc_b200/_native.py).
#   tools/build_variant.sh <name> -DNRP_V_FLUSH_ILP=0 ...
set -eu
name=$1; shift
cd "$(dirname "$0")/../nrpcalc_b200/csrc"
mkdir -p ../../_variants
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wno-unused-function --expt-relaxed-constexpr "$@" \
     -shared -o ../../_variants/libnrpb200_$name.so nrp_b200.cu -lcudart && echo "built $name"
