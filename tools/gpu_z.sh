#!/bin/bash
# round-2 GPU call Z2: k_stats2 without the software pipeline: points per trip x registers x warps per SM (A/B);
# emission after the store / pointer changes
mkdir -p gpurun_out
rm -f gpurun_out/z_kbench.jsonl
for w in "scaleout 32000" "largeN 12500" "gmm100k 100000 3 2"; do
  for lib in "" build/libdnbc_np4.so build/libdnbc_np4r96.so build/libdnbc_np4r80.so build/libdnbc_np4r128w4.so build/libdnbc_np8r96.so; do
    DNBC_B200_LIB=$lib python tools/kbench.py $w >> gpurun_out/z_kbench.jsonl 2>> gpurun_out/z_err.log
  done
done
cat gpurun_out/z_kbench.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'], d['workload'], d['ms_per_iter']['stats_gmm'], d['ms_per_iter']['emission'], d['ll'])"
