#!/bin/bash
# round-2 GPU call L: emission sweep with one log2 per pair of variables -- timings and the GPU suite
mkdir -p gpurun_out
for w in "scaleout 32000" "largeN 12500" "gmm100k 100000"; do python tools/kbench.py $w 3; done > gpurun_out/l_timing.jsonl 2> gpurun_out/l_err.log
python - <<'PY'
import json
for l in open('gpurun_out/l_timing.jsonl'):
    d=json.loads(l); print(d['workload'], d['ms_per_iter'], d['ll'])
PY
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 -k "not two_gpus" > gpurun_out/l_pytest.log 2>&1; tail -15 gpurun_out/l_pytest.log
