#!/bin/bash
# build_variant.sh NAME "-DFLAG=VALUE ..." : builds compute_war_b200/libcomputewar_NAME.so with extra nvcc flags (A/B measurements;
# select it with CW_LIB=libcomputewar_NAME.so; delete it afterwards -- variants bloat the gpurun snapshot)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
cd "$ROOT/compute_war_b200/csrc"
NAME=$1; shift
OBJ="$ROOT/gpurun_out/var_$NAME"
mkdir -p "$OBJ"
for f in api hash_kernels lz4_kernels lzf_kernels fused_kernels decode_kernels pack_kernels digest_kernels synth; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --diag-suppress 128 "$@" -c -o "$OBJ/$f.o" $f.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libcomputewar_$NAME.so "$OBJ"/*.o -lpthread
echo built ../libcomputewar_$NAME.so
