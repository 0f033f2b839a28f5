#!/bin/bash
# ncu evidence for every named workload: launch list of a short bench run + one --set full capture of its BMU kernel.
# usage: tools/profile_all.sh <tag> [workloads...]      (each ncu run follows a plain run of the same command)
tag=$1; shift
mkdir -p gpurun_out
for wl in "${@:-c2}"; do
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --workload $wl"
  timeout 300 $CMD > gpurun_out/${tag}_plain_$wl.log 2>&1 || { echo "$wl plain run failed"; continue; }
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_$wl.csv \
      $CMD > gpurun_out/${tag}_ncu_launches_$wl.log 2>&1
  echo "$wl launch list exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:bmu_ -s 3 -c 1 \
      -o gpurun_out/${tag}_full_$wl -f $CMD > gpurun_out/${tag}_ncu_full_$wl.log 2>&1
  echo "$wl full capture exit $?"
done
