#!/bin/bash
# round-2 GPU call E8: ncu of the N = 64 Viterbi kernels (k_viterbi2 / k_backtrace)
mkdir -p gpurun_out
python tools/dbench.py scaleout 32000 | cut -c1-250
ncu --set full --clock-control none --import-source on -k "regex:k_viterbi2|k_backtrace" -c 2 -o gpurun_out/prof_r02e_vit -f python tools/dbench.py scaleout 8000 > gpurun_out/e8_ncu.log 2>&1; tail -1 gpurun_out/e8_ncu.log
python tools/ncu_summary.py gpurun_out/prof_r02e_vit.ncu-rep --ops > gpurun_out/r02e_ncu_viterbi2_summary.txt; cat gpurun_out/r02e_ncu_viterbi2_summary.txt | cut -c1-400
ncu -i gpurun_out/prof_r02e_vit.ncu-rep --page source --csv --print-source sass | gzip > gpurun_out/r02e_src_vit.csv.gz
