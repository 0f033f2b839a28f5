#!/bin/bash
# ncu evidence for profiles/: (1) launch list of a short bench run, (2) full capture of the top kernel.
# Each ncu run follows a plain run of the SAME command that exited 0 (B200_PROFILING.md).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu ${BENCH_ARGS}"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-bmu_tc} -s 3 -c 2 \
    -o gpurun_out/prof_${TAG:-bmu_tc} -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -n 3 gpurun_out/plain.log gpurun_out/ncu_launches.log gpurun_out/ncu_full.log
