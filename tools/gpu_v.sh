#!/bin/bash
# round-2 GPU call V: issue-rate microbenchmark for the emission bodies; static ring slots in k_fb2 (A/B), tests
mkdir -p gpurun_out
tools/pipe_peaks3 > gpurun_out/v_pipe_peaks3.txt 2>&1; cat gpurun_out/v_pipe_peaks3.txt
for w in "gmm100k 100000 3 2" "readme 16000 3 2" "robot 16000 3 2" "readme 1000 3 2"; do
  for lib in "" build/libdnbc_fb2old.so build/libdnbc_fb2pf2.so build/libdnbc_fb2pf4.so; do
    DNBC_B200_LIB=$lib python tools/kbench.py $w >> gpurun_out/v_kbench.jsonl 2>> gpurun_out/v_err.log
  done
done
cat gpurun_out/v_kbench.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x -k "not two_gpus" > gpurun_out/v_pytest.log 2>&1; tail -3 gpurun_out/v_pytest.log
