#!/bin/bash
# round-2 GPU call X: points-packed emission with cp.async staging -- tests, timings, ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "not two_gpus" > gpurun_out/x_pytest.log 2>&1; tail -3 gpurun_out/x_pytest.log
for w in "scaleout 32000" "largeN 12500" "gmm100k 100000 3 2" "readme 16000 3 2" "robot 16000 3 2"; do
  for lib in "" build/libdnbc_pxall.so; do
    DNBC_B200_LIB=$lib python tools/kbench.py $w >> gpurun_out/x_kbench.jsonl 2>> gpurun_out/x_err.log
  done
done
cat gpurun_out/x_kbench.jsonl
ncu --set full --clock-control none --import-source on -k "regex:k_emission" -c 1 -o gpurun_out/prof_r02x_px512 -f python tools/kbench.py largeN 12500 1 > gpurun_out/x_ncu1.log 2>&1; tail -1 gpurun_out/x_ncu1.log
DNBC_B200_LIB=build/libdnbc_pxall.so ncu --set full --clock-control none --import-source on -k "regex:k_emission" -c 1 -o gpurun_out/prof_r02x_px64 -f python tools/kbench.py scaleout 8000 1 > gpurun_out/x_ncu2.log 2>&1; tail -1 gpurun_out/x_ncu2.log
python tools/ncu_summary.py gpurun_out/prof_r02x_px512.ncu-rep --ops > gpurun_out/r02x_ncu_px512_summary.txt
python tools/ncu_summary.py gpurun_out/prof_r02x_px64.ncu-rep --ops > gpurun_out/r02x_ncu_px64_summary.txt
for r in px512 px64; do ncu -i gpurun_out/prof_r02x_$r.ncu-rep --page source --csv --print-source sass | gzip > gpurun_out/r02x_src_$r.csv.gz; done
cat gpurun_out/r02x_ncu_px512_summary.txt gpurun_out/r02x_ncu_px64_summary.txt
find gpurun_out -name "*.ncu-rep" -size +20M -delete
