#!/bin/bash
# usage: build_variants.sh name "flags" [name "flags"]...
cd "$(dirname "$0")"
rm -f k1e_v_*
while [ $# -gt 1 ]; do
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I../../covariancefunctions.jl_b200/csrc -Xcudafe --diag_suppress=177 $2 -o k1e_v_$1 k1e_variants.cu &
  shift 2
done
wait
ls k1e_v_*
