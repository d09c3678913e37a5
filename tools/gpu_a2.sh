#!/bin/bash
# round-2 GPU call A2: k_stats2 without the software pipeline and with the points-packed mode for odd mixtures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "not two_gpus" > gpurun_out/a2_pytest.log 2>&1; tail -3 gpurun_out/a2_pytest.log
for w in "scaleout 32000" "largeN 12500" "gmm100k 100000 3 2" "readme 16000 3 2" "robot 16000 3 2"; do
  for lib in "" build/libdnbc_s2nopx.so; do
    DNBC_B200_LIB=$lib python tools/kbench.py $w >> gpurun_out/a2_kbench.jsonl 2>> gpurun_out/a2_err.log
  done
done
cat gpurun_out/a2_kbench.jsonl
