#!/bin/bash
# usage: tools/sweep_env.sh VAR "v1 v2 ..." "workload seqs" ...   -- kbench totals per value of an environment switch
var=$1; vals=$2; shift 2
for v in $vals; do
  for w in "$@"; do
    env $var=$v python tools/kbench.py $w 3 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$var=$v', d['workload'], d['seqs'], {k: d['ms_per_iter'][k] for k in d['ms_per_iter']})"
  done
done
