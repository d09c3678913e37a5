#!/bin/bash
# round-2 GPU call C2: statistics sweep old / new at the bench size (125,000 sequences), alternating
mkdir -p gpurun_out
for rep in 1 2; do
  for lib in "" build/libdnbc_s2old.so; do
    DNBC_B200_LIB=$lib python tools/kbench.py scaleout 125000 3 >> gpurun_out/c2_kbench.jsonl 2>> gpurun_out/c2_err.log
  done
done
cat gpurun_out/c2_kbench.jsonl
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,temperature.gpu --format=csv
