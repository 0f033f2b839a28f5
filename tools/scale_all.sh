#!/bin/bash
# the default bench line (c2 + sub-records) on N GPUs of one box, as the driver launches it
N=$1; mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "N=$N exit $?"
