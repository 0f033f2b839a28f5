#!/bin/bash
# One gpurun call: bring-up diagnostics, GPU tests, bench.  Every stage has its own timeout so a
# faulting kernel cannot take the whole call (or the box) with it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
for st in simt tc epoch; do
  timeout 300 python tools/first_light.py $st > gpurun_out/first_light_$st.log 2>&1
  echo "first_light $st exit $?" >> gpurun_out/summary.log
done
timeout 900 python -m pytest tests -m gpu -q -x -k "not tc" > gpurun_out/pytest_simt.log 2>&1
echo "pytest not-tc exit $?" >> gpurun_out/summary.log
timeout 900 python -m pytest tests -m gpu -q -k "tc" > gpurun_out/pytest_tc.log 2>&1
echo "pytest tc exit $?" >> gpurun_out/summary.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log
tail -n 30 gpurun_out/first_light_*.log
tail -n 5 gpurun_out/pytest_simt.log gpurun_out/pytest_tc.log gpurun_out/bench.log
