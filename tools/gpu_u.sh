#!/bin/bash
# round-2 GPU call U: points-packed emission kernel (k_emission_px) and scalar-broadcast observation in k_stats2:
# GPU suite, then A/B timings against the previous kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "not two_gpus" > gpurun_out/u_pytest.log 2>&1; tail -5 gpurun_out/u_pytest.log
for w in "scaleout 32000" "largeN 12500" "gmm100k 100000 3 2" "readme 16000 3 2" "robot 16000 3 2"; do
  for lib in "" build/libdnbc_emold.so build/libdnbc_s2old.so build/libdnbc_pxall.so; do
    DNBC_B200_LIB=$lib python tools/kbench.py $w >> gpurun_out/u_kbench.jsonl 2>> gpurun_out/u_err.log
  done
done
cat gpurun_out/u_kbench.jsonl
