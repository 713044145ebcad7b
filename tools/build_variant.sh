#!/bin/bash
# build_variant.sh NAME "-DFLAG=VALUE ..." : builds compute_war_b200/libcomputewar_NAME.so with extra nvcc flags (A/B measurements)
set -e
cd "$(dirname "$0")/../compute_war_b200/csrc"
NAME=$1; shift
mkdir -p /tmp/cwvar_$NAME
for f in api hash_kernels lz4_kernels lzf_kernels fused_kernels decode_kernels pack_kernels synth; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --diag-suppress 128 "$@" -c -o /tmp/cwvar_$NAME/$f.o $f.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libcomputewar_$NAME.so /tmp/cwvar_$NAME/*.o -lpthread
echo built ../libcomputewar_$NAME.so
