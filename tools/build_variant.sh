#!/bin/bash
# usage: tools/build_variant.sh out.so [-DB32_OP_THREADS=128 -DB32_OP_DUAL=0 ...]   (experiment builds; run with B32_LIB=out.so)
set -e
out=$1; shift
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC,-O2,-ffp-contract=off -shared -cudart static "$@" -I include -o "$out" bonnie-32_b200/csrc/*.cu
