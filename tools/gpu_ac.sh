#!/bin/bash
# round-2 GPU call AC: k_xi_tc with free off-items (partial tiles): parity tests and timings at in-between state counts
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x  > gpurun_out/ac_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/ac_pytest.log; tail -3 gpurun_out/ac_pytest.log
rm -f gpurun_out/ac_kbench.jsonl
for w in "custom:N=100,T=200 50000" "custom:N=200,T=200 25000" "custom:N=300,T=200 12500" "custom:N=320,T=200 12500" "custom:N=500,T=200 12500" "largeN 12500" "scaleout 8000"; do
  timeout 300 python tools/kbench.py $w 2 >> gpurun_out/ac_kbench.jsonl 2>> gpurun_out/ac_err.log
done
python - <<'PY'
import json
for l in open("gpurun_out/ac_kbench.jsonl"):
    d = json.loads(l); m = d["ms_per_iter"]
    print(d["lib"], d["workload"], d["seqs"], "fwd", m.get("forward"), "bwd", m.get("backward"), "xi", m.get("xi"), "em", m.get("emission"), "st", m.get("stats_gmm"), "total", m["total"], "ll", d["ll"])
PY
tail -5 gpurun_out/ac_err.log 2>/dev/null
