#!/bin/bash
# A/B builds: tools/build_variant.sh <name> <file.cu> "<extra nvcc flags>"  ->  build/libdnbc_<name>.so
# (one translation unit recompiled with the flags, the rest taken from dnbc_scala_b200/lib/*.o)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; src=$2; flags=$3
mkdir -p $ROOT/build
stem=$(basename $src .cu)
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr $flags \
  -c -o $ROOT/build/${stem}_$name.o $ROOT/dnbc_scala_b200/csrc/$src
objs=$(ls $ROOT/dnbc_scala_b200/lib/*.o | grep -v "/$stem.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/build/libdnbc_$name.so $objs $ROOT/build/${stem}_$name.o -lcudart -lnccl
echo built build/libdnbc_$name.so
