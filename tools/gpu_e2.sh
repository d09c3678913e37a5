#!/bin/bash
# round-2 GPU call E6: tile depth of the lane-per-sequence recursions (4 / 2 steps) forward and backward
mkdir -p gpurun_out
rm -f gpurun_out/e6_kbench.jsonl
for w in "gmm100k 100000 3 2" "robot 100000 3 2" "gmm100k 25000 3 2"; do
for lib in "" build/libdnbc_tbf2.so build/libdnbc_tbb2.so build/libdnbc_tb22.so; do
  DNBC_FB_SEQ_LANES=1 DNBC_B200_LIB=$lib python tools/kbench.py $w >> gpurun_out/e6_kbench.jsonl 2>> gpurun_out/e6_err.log
done
done
python - <<'PY'
import json
for l in open('gpurun_out/e6_kbench.jsonl'):
    d = json.loads(l); m = d['ms_per_iter']; print(d['lib'], d['workload'], d['seqs'], 'fwd', m['forward'], 'bwd', m['backward'], d['ll'])
PY
