#!/bin/bash
# round-2 GPU call E3: k_stats2 with 8-byte observation records against the previous commit
mkdir -p gpurun_out
rm -f gpurun_out/e2_kbench.jsonl
for w in "scaleout 32000 3" "largeN 12500 3" "gmm100k 100000 3 2"; do
for lib in "" build/libdnbc_s2prev.so "" build/libdnbc_s2prev.so; do
  DNBC_B200_LIB=$lib python tools/kbench.py $w >> gpurun_out/e2_kbench.jsonl 2>> gpurun_out/e2_err.log
done
done
python - <<'PY'
import json
for l in open('gpurun_out/e2_kbench.jsonl'):
    d = json.loads(l); print(d['lib'], d['workload'], d['ms_per_iter']['stats_gmm'], d['ll'])
PY
timeout 1500 python -m pytest tests -m gpu -q -x -k "not two_gpus" > gpurun_out/e2_pytest.log 2>&1; tail -3 gpurun_out/e2_pytest.log
