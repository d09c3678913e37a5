#!/bin/bash
# round-2 GPU call E7: k_stats2 with the odd last component of two points as one packed chain, against the previous commit
mkdir -p gpurun_out
rm -f gpurun_out/e7_kbench.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -k "estep or stats or config5 or fused_and_general or randomised or outlier or lane_per" > gpurun_out/e7_pytest.log 2>&1; tail -3 gpurun_out/e7_pytest.log
for w in "largeN 12500 3" "gmm100k 100000 3 2" "readme 16000 3 2"; do
for lib in "" build/libdnbc_s2prev.so "" build/libdnbc_s2prev.so; do
  DNBC_B200_LIB=$lib python tools/kbench.py $w >> gpurun_out/e7_kbench.jsonl 2>> gpurun_out/e7_err.log
done
done
python - <<'PY'
import json
for l in open('gpurun_out/e7_kbench.jsonl'):
    d = json.loads(l); print(d['lib'], d['workload'], d['ms_per_iter']['stats_gmm'], d['ll'])
PY
