#!/bin/bash
# time every library variant under tools/variants on the named shapes (BMU only and fused)
for lib in tools/variants/*.so; do
  echo "=== $lib"
  for shape in "2000000 16 1600" "1000000 64 1024" "500000 128 2500"; do
    for mode in nofuse fused; do
      SOM_TOOL_LIB=$lib timeout 60 python tools/bmu_probe.py $shape $mode 5 tc16 2>&1 | grep -v "^tile\|^ *[0-9]* |" | tail -1
    done
  done
done
