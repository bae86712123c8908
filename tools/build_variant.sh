#!/bin/bash
# Build a second copy of the library with extra nvcc defines (A/B timing of compile-time switches):
#   tools/build_variant.sh lane0 -DIVV_ROLE_ELECT=0   ->  insv2v_b200/libivv_b200_lane0.so   (use with IVV_LIB_PATH)
set -e
name=$1; shift
cd "$(dirname "$0")/../insv2v_b200"
mkdir -p build_$name
for f in api gemm_tc attention_tc norm temporal_attn elementwise warp sampler raft; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
       -c csrc/$f.cu -o build_$name/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libivv_b200_$name.so build_$name/*.o -lcudart
echo built libivv_b200_$name.so
