#!/bin/bash
# Build variants of the MaternP scaled-domain kernel (gram_mvm_eq.cuh FAST = 2: CF_MVMM_R rows per thread at D <= 4, CF_MVMM_R8 at D = 8)
# as complete libraries under bench_aux/micro/variants/:   k1m_variants.sh name "flags" [name "flags"] ...
cd "$(dirname "$0")/../../covariancefunctions.jl_b200/csrc"
B=../../build/covfn
V=../../bench_aux/micro/variants
mkdir -p $V
while [ $# -gt 1 ]; do
  ( for d in 3 8; do nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O2,-Wall -Xcudafe --diag_suppress=177 -DCF_D=$d $2 -c cf_inst.cu -o $V/cf_inst_d${d}_$1.o; done
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $V/libcovfn_$1.so $B/capi.o $(for d in 1 2 4 6 12 16 24 32; do echo $B/cf_inst_d$d.o; done) $V/cf_inst_d3_$1.o $V/cf_inst_d8_$1.o -lcudart -ldl; rm -f $V/*_$1.o ) &
  shift 2
done
wait
ls $V/*.so
