#!/bin/bash
# round-2 GPU call M: k_stats2 in its grouped layout (N <= 32) -- GPU suite and small-state timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 -k "not two_gpus" > gpurun_out/m_pytest.log 2>&1; tail -15 gpurun_out/m_pytest.log
for w in "gmm100k 100000 3 2" "readme 16000 3 2" "robot 16000 3 2" "scaleout 32000 3" "largeN 12500 3"; do python tools/kbench.py $w; done > gpurun_out/m_timing.jsonl 2> gpurun_out/m_err.log
python - <<'PY'
import json
for l in open('gpurun_out/m_timing.jsonl'):
    d=json.loads(l); print(d['workload'], d['seqs'], d['ms_per_iter'])
PY
