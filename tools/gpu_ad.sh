#!/bin/bash
# round-2 GPU call AD: k_stats2 with the partial last slice against the previous build at the BASELINE shapes (same call)
mkdir -p gpurun_out
rm -f gpurun_out/ad_kbench.jsonl
for rep in 1 2; do
for w in "scaleout 32000" "largeN 12500"; do
  for lib in "" build/libdnbc_s2old.so; do
    DNBC_B200_LIB=$lib timeout 300 python tools/kbench.py $w 3 >> gpurun_out/ad_kbench.jsonl 2>> gpurun_out/ad_err.log
  done
done
done
python - <<'PY'
import json
for l in open("gpurun_out/ad_kbench.jsonl"):
    d = json.loads(l); m = d["ms_per_iter"]
    print(d["lib"], d["workload"], d["seqs"], "em", m.get("emission"), "st", m.get("stats_gmm"), "total", m["total"], "ll", d["ll"])
PY
tail -5 gpurun_out/ad_err.log 2>/dev/null
