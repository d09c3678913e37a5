#!/bin/bash
# A/B builds from an alternative source file: tools/build_variant_src.sh <name> <path/to/file.cu> "<extra nvcc flags>"
# -> build/libdnbc_<name>.so (that translation unit replaces the one of the same name; the rest from dnbc_scala_b200/lib/*.o)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; src=$2; flags=$3
mkdir -p $ROOT/build
stem=$(basename $src .cu)
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr $flags \
  -I$ROOT/dnbc_scala_b200/csrc -c -o $ROOT/build/${stem}_$name.o $src
objs=$(ls $ROOT/dnbc_scala_b200/lib/*.o | grep -v "/$stem.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/build/libdnbc_$name.so $objs $ROOT/build/${stem}_$name.o -lcudart -lnccl
echo built build/libdnbc_$name.so
