#!/bin/bash
# Weak-scaling bench lines on N GPUs of one box (N = first argument), one line per workload.
N=${1:-8}
mkdir -p gpurun_out
for w in ${WORKLOADS:-c2 c3 c4}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_${w}_n$N.log 2>&1
  echo "$w N=$N exit $?"
  tail -n 1 gpurun_out/scale_${w}_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('  value %.3e  ms/step %.3f  kernel_ms %.3f  e2e %.3e' % (d['value'], d['ms_per_step'], r['kernel_ms'], d['e2e']['value']))"
done
