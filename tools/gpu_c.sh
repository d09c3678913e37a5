#!/bin/bash
# round-2 GPU call C: ncu --set full of the statistics and emission sweeps (8e6 points per launch)
mkdir -p gpurun_out
CMD="python tools/kbench.py scaleout 8000 1"
$CMD > gpurun_out/c_plain.log 2>&1 || { tail -5 gpurun_out/c_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:k_stats2|k_emission_pk" -s 4 -c 2 -o gpurun_out/prof_r02c -f $CMD > gpurun_out/c_ncu.log 2>&1
tail -3 gpurun_out/c_ncu.log
ls -la gpurun_out/prof_r02c.ncu-rep
