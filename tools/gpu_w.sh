#!/bin/bash
# round-2 GPU call W: why is the points-packed emission kernel slower at N = 512 / K = 8?  microbenchmark + ncu
mkdir -p gpurun_out
timeout 120 tools/pipe_peaks3 > gpurun_out/w_pipe_peaks3.txt 2>&1; echo "pipe_peaks3 rc=$?"; cat gpurun_out/w_pipe_peaks3.txt
python tools/kbench.py gmm100k 100000 3 2 > gpurun_out/w_kbench.jsonl 2>> gpurun_out/w_err.log; cat gpurun_out/w_kbench.jsonl
ncu --set full --clock-control none --import-source on -k "regex:k_emission" -c 1 -o gpurun_out/prof_r02w_px512 -f python tools/kbench.py largeN 12500 1 > gpurun_out/w_ncu1.log 2>&1; tail -1 gpurun_out/w_ncu1.log
DNBC_B200_LIB=build/libdnbc_pxall.so ncu --set full --clock-control none --import-source on -k "regex:k_emission" -c 1 -o gpurun_out/prof_r02w_px64 -f python tools/kbench.py scaleout 8000 1 > gpurun_out/w_ncu2.log 2>&1; tail -1 gpurun_out/w_ncu2.log
python tools/ncu_summary.py gpurun_out/prof_r02w_px512.ncu-rep --ops > gpurun_out/r02w_ncu_px512_summary.txt
python tools/ncu_summary.py gpurun_out/prof_r02w_px64.ncu-rep --ops > gpurun_out/r02w_ncu_px64_summary.txt
for r in px512 px64; do ncu -i gpurun_out/prof_r02w_$r.ncu-rep --page source --csv --print-source sass | gzip > gpurun_out/r02w_src_$r.csv.gz; done
cat gpurun_out/r02w_ncu_px512_summary.txt gpurun_out/r02w_ncu_px64_summary.txt
find gpurun_out -name "*.ncu-rep" -size +20M -delete
