#!/bin/bash
# quick tuning variant: only ONE translation unit (default csrc/hpe_cost.cu, the headline kernel's; SRC=hpe_obs for another)
# is recompiled with the given -D flags and linked with the standard objects.
# tools/build_cost_variant.sh NAME -DHPE_X=.. ; select with HPE_B200_LIB
set -e
cd "$(dirname "$0")/../hand-pose-estimation_b200"
name=$1; shift
src=${SRC:-hpe_cost}
mkdir -p build/variants/$name
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
   -ccbin /usr/bin/g++ "$@" -c csrc/$src.cu -o build/variants/$name/$src.o
objs=$(ls build/*.o | grep -v "/$src\.o$")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static \
   -o build/variants/libhpe_b200_$name.so build/variants/$name/$src.o $objs
echo built build/variants/libhpe_b200_$name.so
