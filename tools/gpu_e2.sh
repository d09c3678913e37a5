#!/bin/bash
# round-2 GPU call E11: deferred exponent-shifted re-evaluations in k_emission_px against the previous commit
mkdir -p gpurun_out
rm -f gpurun_out/e11_kbench.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -k "emission or outlier or config5 or n512 or randomised or fused_and_general" > gpurun_out/e11_pytest.log 2>&1; tail -3 gpurun_out/e11_pytest.log
for w in "largeN 12500 3" "gmm100k 100000 3 2" "readme 16000 3 2"; do
for lib in "" build/libdnbc_pxprev.so "" build/libdnbc_pxprev.so; do
  DNBC_B200_LIB=$lib python tools/kbench.py $w >> gpurun_out/e11_kbench.jsonl 2>> gpurun_out/e11_err.log
done
done
python - <<'PY'
import json
for l in open('gpurun_out/e11_kbench.jsonl'):
    d = json.loads(l); print(d['lib'], d['workload'], d['ms_per_iter']['emission'], d['ll'])
PY
