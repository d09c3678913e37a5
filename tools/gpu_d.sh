#!/bin/bash
# round-2 GPU call D: emission sweep with 0 / 1 / 2 component pairs per variable evaluated by ex2_poly2
mkdir -p gpurun_out
: > gpurun_out/d_variants.jsonl
for v in ep0 ep1 ep2; do
  DNBC_B200_LIB=build/libdnbc_$v.so python tools/kbench.py scaleout 32000 3 >> gpurun_out/d_variants.jsonl 2>> gpurun_out/d_err.log
done
cat gpurun_out/d_variants.jsonl
