#!/bin/bash
# bench lines for every named workload (one GPU) + launch lists for c4
mkdir -p gpurun_out
for w in c2 c3 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$w.log 2>&1
  echo "$w exit $?"; tail -n 1 gpurun_out/bench_$w.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('  value %.3e  ms/step %.3f  kernel_ms %.3f  achieved %.1f TF  frac %.3f  e2e %.3e' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['achieved'], r['frac'] or 0, d['e2e']['value']))"
done
CMD="python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c4.csv $CMD > gpurun_out/ncu_launches_c4.log 2>&1
echo "c4 launch list exit $?"
