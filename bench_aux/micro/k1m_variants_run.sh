#!/bin/bash
# on the GPU box: MaternP(2) value MVMs (d = 3 at n = 16384 and 131072, d = 8 at n = 131072) with every library under bench_aux/micro/variants/
cd "$(dirname "$0")/../.."
L=covariancefunctions.jl_b200/lib/libcovfn_b200.so
cp $L /tmp/libcovfn_default.so
for v in /tmp/libcovfn_default.so bench_aux/micro/variants/libcovfn_*.so; do
  cp $v $L 2>/dev/null
  echo "$v c1 $(python bench_aux/run_one.py --config c1 --reps 5 | cut -c40-75) | c1@131072 $(python bench_aux/run_one.py --config c1 --n 131072 --reps 3 | cut -c40-78) | x4 $(python bench_aux/run_one.py --config x4 --reps 3 | cut -c40-78)"
done
cp /tmp/libcovfn_default.so $L
