#!/bin/bash
# time the production library's BMU kernel (search only, and fused with the accumulate) on the named shapes
for shape in "2000000 16 1600" "1000000 64 1024" "500000 128 2500" ${EXTRA_SHAPES}; do
  for mode in nofuse fused; do
    timeout 60 python tools/bmu_probe.py $shape $mode 5 ${ALGO:-tc16} 2>&1 | tail -1
  done
done
