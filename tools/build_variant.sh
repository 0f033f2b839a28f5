#!/bin/bash
# usage: tools/build_variant.sh <name> [-DFLAG ...]   -> tools/variants/lib_<name>.so (experiments build of the library)
name=$1; shift
mkdir -p tools/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -cudart static \
  -DSOM_B200_EXPERIMENTS "$@" -o tools/variants/lib_$name.so xpysom_dask_b200/csrc/som_api.cu
