#!/bin/bash
# round-2 GPU call Y: points-packed emission (cheap request) for every shape; statistics-body microbenchmark
mkdir -p gpurun_out
timeout 120 tools/pipe_peaks3 > gpurun_out/y_pipe_peaks3.txt 2>&1; tail -3 gpurun_out/y_pipe_peaks3.txt
timeout 1500 python -m pytest tests -m gpu -q -x -k "not two_gpus" > gpurun_out/y_pytest.log 2>&1; tail -3 gpurun_out/y_pytest.log
for w in "scaleout 32000" "largeN 12500" "gmm100k 100000 3 2" "readme 16000 3 2" "robot 16000 3 2"; do
  for lib in "" build/libdnbc_pxnone.so; do
    DNBC_B200_LIB=$lib python tools/kbench.py $w >> gpurun_out/y_kbench.jsonl 2>> gpurun_out/y_err.log
  done
done
cat gpurun_out/y_kbench.jsonl
ncu --set full --clock-control none --import-source on -k "regex:k_emission" -c 1 -o gpurun_out/prof_r02y_px512 -f python tools/kbench.py largeN 12500 1 > gpurun_out/y_ncu1.log 2>&1; tail -1 gpurun_out/y_ncu1.log
python tools/ncu_summary.py gpurun_out/prof_r02y_px512.ncu-rep --ops > gpurun_out/r02y_ncu_px512_summary.txt
ncu -i gpurun_out/prof_r02y_px512.ncu-rep --page source --csv --print-source sass | gzip > gpurun_out/r02y_src_px512.csv.gz
cat gpurun_out/r02y_ncu_px512_summary.txt
find gpurun_out -name "*.ncu-rep" -size +20M -delete
