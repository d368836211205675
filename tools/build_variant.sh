#!/bin/bash
# build a tuning variant of libhpe_b200.so: tools/build_variant.sh NAME -DHPE_X=.. ...
# result: hand-pose-estimation_b200/build/variants/libhpe_b200_NAME.so (select with HPE_B200_LIB)
set -e
cd "$(dirname "$0")/../hand-pose-estimation_b200"
name=$1; shift
mkdir -p build/variants/$name
for f in hpe_api hpe_cost hpe_obs hpe_pso hpe_pso_cluster hpe_swarm hpe_refine hpe_pso_opt hpe_bincomp hpe_depthmatch hpe_bench; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
     -ccbin /usr/bin/g++ "$@" -c csrc/$f.cu -o build/variants/$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static \
   -o build/variants/libhpe_b200_$name.so build/variants/$name/*.o
echo built build/variants/libhpe_b200_$name.so
