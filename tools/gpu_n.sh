#!/bin/bash
mkdir -p gpurun_out
for w in "scaleout 32000 3" "largeN 12500 3" "gmm100k 100000 3 2"; do python tools/kbench.py $w; done > gpurun_out/n_timing.jsonl 2> gpurun_out/n_err.log
python - <<'PY'
import json
for l in open('gpurun_out/n_timing.jsonl'):
    d=json.loads(l); print(d['workload'], d['seqs'], d['ms_per_iter'])
PY
