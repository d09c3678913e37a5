#!/bin/bash
# round-2 GPU call AI: E-step chunks cut at whole waves of 128-sequence batches (tcgen05 recursions)
mkdir -p gpurun_out
rm -f gpurun_out/ai_kbench.jsonl
for w in "largeN 100000" "largeN 50000" "custom:N=300,T=200 100000" "scaleout 32000"; do
  timeout 300 python tools/kbench.py $w 2 >> gpurun_out/ai_kbench.jsonl 2>> gpurun_out/ai_err.log
done
python - <<'PY'
import json
for l in open("gpurun_out/ai_kbench.jsonl"):
    d = json.loads(l); m = d["ms_per_iter"]
    print(d["lib"], d["workload"], d["seqs"], "em", m.get("emission"), "fwd", m.get("forward"), "bwd", m.get("backward"), "xi", m.get("xi"), "st", m.get("stats_gmm"), "total", m["total"], "ll", d["ll"])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "chunk or tensor_core" > gpurun_out/ai_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ai_pytest.log; tail -2 gpurun_out/ai_pytest.log
tail -5 gpurun_out/ai_err.log 2>/dev/null
