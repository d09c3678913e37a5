#!/bin/bash
# round-2 GPU call H: fused vs general E-step over the number of sequences (README shape and config-3 shape)
mkdir -p gpurun_out
: > gpurun_out/h_crossover.jsonl
for w in readme gmm100k robot; do for s in 250 500 1000 2000 4000 8000 16000; do for m in 1 2; do
  python tools/kbench.py $w $s 5 $m 2>> gpurun_out/h_err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); d['fused_mode']=$m; print(json.dumps(d))" >> gpurun_out/h_crossover.jsonl
done; done; done
python - <<'PY'
import json
for l in open('gpurun_out/h_crossover.jsonl'):
    d=json.loads(l); print(d['workload'], d['seqs'], 'fused' if d['fused_mode']==1 else 'general', d['ms_per_iter']['total'])
PY
