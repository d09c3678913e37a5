#!/bin/bash
# round-2 GPU call O: ncu of the max-plus Viterbi tile kernel at N = 512 (and the GPU suite after the dispatch change)
mkdir -p gpurun_out
CMD="python tools/dbench.py largeN 12000"
$CMD > gpurun_out/o_plain.log 2>&1 || { tail -5 gpurun_out/o_plain.log; exit 1; }
cat gpurun_out/o_plain.log
ncu --set full --clock-control none --import-source on -k "regex:k_viterbi_tile" -c 1 -o gpurun_out/prof_r02o -f $CMD > gpurun_out/o_ncu.log 2>&1
tail -2 gpurun_out/o_ncu.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 -k "not two_gpus" > gpurun_out/o_pytest.log 2>&1; tail -4 gpurun_out/o_pytest.log
