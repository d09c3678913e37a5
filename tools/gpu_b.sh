#!/bin/bash
# round-2 GPU call B: statistics-sweep variants (points per loop trip) against the round-1 kernel
mkdir -p gpurun_out build
# (variant libraries are built in the dev container with tools/build_variant.sh and travel with the snapshot)
: > gpurun_out/b_variants.jsonl
for v in nos2 s2p4 s2p8 s2p16 s2p32; do
  DNBC_B200_LIB=build/libdnbc_$v.so python tools/kbench.py scaleout 32000 3 >> gpurun_out/b_variants.jsonl 2>> gpurun_out/b_err.log
  DNBC_B200_LIB=build/libdnbc_$v.so python tools/kbench.py largeN 12500 3 >> gpurun_out/b_variants.jsonl 2>> gpurun_out/b_err.log
done
cat gpurun_out/b_variants.jsonl
