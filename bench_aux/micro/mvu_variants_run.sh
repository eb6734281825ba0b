#!/bin/bash
# on the GPU box: time the auxiliary Float32 EQ MVMs (d = 32 and d = 8, n = 131072) with every library under bench_aux/micro/variants/
cd "$(dirname "$0")/../.."
L=covariancefunctions.jl_b200/lib/libcovfn_b200.so
cp $L /tmp/libcovfn_default.so
for v in bench_aux/micro/variants/libcovfn_*.so; do
  cp $v $L
  for c in x2 x7; do echo "$v $c $(python bench_aux/run_one.py --config $c --dtype f32 --reps 4 | cut -c1-80)"; done
done
cp /tmp/libcovfn_default.so $L
