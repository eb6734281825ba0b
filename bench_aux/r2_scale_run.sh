#!/bin/bash
# Round-2 multi-GPU evidence run (gpurun --gpus 8): multi-device tests, then BASELINE configs 2 and 5 at N = 2, 4, 8 under torchrun
# (one process per GPU, NCCL inside the library) and, for config 5, in the single-process mode (cf_init, peer loads / stores).
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n8.txt 2>&1
python -m pytest tests/test_gpu_multidevice.py -v -m gpu 2>&1 | tail -14 > gpurun_out/r2_multidevice_n8.log
P=29600
for N in 2 4 8; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --config c2 --steps 3 > gpurun_out/r2_bench_c2_n$N.json 2> gpurun_out/r2_bench_c2_n$N.err
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --config c5 --steps 1 > gpurun_out/r2_bench_c5_n$N.json 2> gpurun_out/r2_bench_c5_n$N.err
  python bench.py --gpus $N --config c5 --steps 1 --spmd > gpurun_out/r2_bench_c5_spmd_n$N.json 2> gpurun_out/r2_bench_c5_spmd_n$N.err
done
tail -n 3 gpurun_out/r2_multidevice_n8.log
for f in gpurun_out/r2_bench_c*_n[248].json; do echo "$f: $(tail -c 200 $f)"; done
