#!/bin/bash
# round-2 GPU call G: ncu of the fused small-N E-step (config 3 shape, 30,000 sequences)
mkdir -p gpurun_out
CMD="python tools/kbench.py gmm100k 30000 1"
$CMD > gpurun_out/g_plain.log 2>&1 || { tail -5 gpurun_out/g_plain.log; exit 1; }
cat gpurun_out/g_plain.log
ncu --set full --clock-control none --import-source on -k "regex:k_estep_fused" -s 2 -c 1 -o gpurun_out/prof_r02g -f $CMD > gpurun_out/g_ncu.log 2>&1
tail -2 gpurun_out/g_ncu.log
