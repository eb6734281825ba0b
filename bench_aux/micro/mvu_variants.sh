#!/bin/bash
# Build variants of the tcgen05 Float32 value MVM (gram_mvm_tc5.cuh macros) as complete libraries under bench_aux/micro/variants/:
#   mvu_variants.sh name "flags" [name "flags"] ...       e.g.  p3 "-DCF_MVU_POLY_MASK=0x1084u"
cd "$(dirname "$0")/../../covariancefunctions.jl_b200/csrc"
B=../../build/covfn
V=../../bench_aux/micro/variants
mkdir -p $V
while [ $# -gt 1 ]; do
  ( for d in 8 32; do nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O2,-Wall -Xcudafe --diag_suppress=177 -DCF_D=$d $2 -c cf_inst.cu -o $V/cf_inst_d${d}_$1.o; done
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $V/libcovfn_$1.so $B/capi.o $(for d in 1 2 3 4 6 12 16 24; do echo $B/cf_inst_d$d.o; done) $V/cf_inst_d8_$1.o $V/cf_inst_d32_$1.o -lcudart -ldl; rm -f $V/*_$1.o ) &
  shift 2
done
wait
ls $V/*.so
