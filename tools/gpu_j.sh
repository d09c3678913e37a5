#!/bin/bash
# round-2 GPU call J: sensitivity of k_stats2 to its FMA work, its histogram and its MUFU work (wrong results on purpose)
mkdir -p gpurun_out
: > gpurun_out/j_sens.jsonl
for v in xbase xdrops2 xdropcat xlessmufu xboth; do
  DNBC_B200_LIB=build/libdnbc_$v.so python tools/kbench.py scaleout 32000 3 >> gpurun_out/j_sens.jsonl 2>> gpurun_out/j_err.log
done
python - <<'PY'
import json
for l in open('gpurun_out/j_sens.jsonl'):
    d=json.loads(l); print(d['lib'], d['ms_per_iter'].get('stats_gmm'))
PY
