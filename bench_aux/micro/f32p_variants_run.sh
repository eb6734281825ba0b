#!/bin/bash
# on the GPU box: time config 2 in Float32 with every library under build/variants/
cd "$(dirname "$0")/../.."
L=covariancefunctions.jl_b200/lib/libcovfn_b200.so
cp $L /tmp/libcovfn_default.so
for v in bench_aux/micro/variants/libcovfn_*.so; do
  cp $v $L
  python bench.py --config c2 --dtype f32 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['roofline']['kernel_ms'],2), 'ms', d['parity_check']['ok'])"
done
cp /tmp/libcovfn_default.so $L
