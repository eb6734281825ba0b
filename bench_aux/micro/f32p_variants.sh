#!/bin/bash
# Build variants of the packed Float32 kernel (gram_mvm_f32p.cuh tuning macros) as complete libraries under build/variants/:
#   f32p_variants.sh name "flags" [name "flags"] ...       e.g.  r8 "-DCF_MVP_R=8 -DCF_MVP_NT=128 -DCF_MVP_MINB=4"
# and run them on the GPU box with bench_aux/micro/f32p_variants_run.sh (copies each over lib/libcovfn_b200.so, restores the default).
cd "$(dirname "$0")/../../covariancefunctions.jl_b200/csrc"
B=../../build/covfn
V=../../bench_aux/micro/variants
mkdir -p $V
while [ $# -gt 1 ]; do
  ( nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O2,-Wall -Xcudafe --diag_suppress=177 -DCF_D=3 $2 -c cf_inst.cu -o $V/cf_inst_d3_$1.o &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $V/libcovfn_$1.so $B/capi.o $(for d in 1 2 4 6 8 12 16 24 32; do echo $B/cf_inst_d$d.o; done) $V/cf_inst_d3_$1.o -lcudart -ldl &&
    cuobjdump -res-usage $V/cf_inst_d3_$1.o 2>/dev/null | grep -A1 "f32p_kernelILi3ELi0" | grep -o "REG:[0-9]*" | sed "s/^/$1 /" ) &
  shift 2
done
wait
ls $V/*.so
