#!/bin/bash
# round-2 GPU call AG: k_xi_tc with L2 requests two K blocks ahead against the plain kernel (same call)
mkdir -p gpurun_out
rm -f gpurun_out/ag_kbench.jsonl
for w in "largeN 12500" "largeN 100000" "custom:N=128,T=200 50000" "custom:N=300,T=200 12500"; do
  for lib in "" build/libdnbc_xipf.so; do
    DNBC_B200_LIB=$lib timeout 300 python tools/kbench.py $w 2 >> gpurun_out/ag_kbench.jsonl 2>> gpurun_out/ag_err.log
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/ag_kbench.jsonl"):
    d = json.loads(l); m = d["ms_per_iter"]
    print(d["lib"], d["workload"], d["seqs"], "fwd", m.get("forward"), "bwd", m.get("backward"), "xi", m.get("xi"), "total", m["total"], "ll", d["ll"])
PY
tail -5 gpurun_out/ag_err.log 2>/dev/null
