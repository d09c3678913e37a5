#!/bin/bash
# round-2 GPU call K: second moment through the exponent argument -- timing, parity tests, parameter deviation
mkdir -p gpurun_out
for v in "" build/libdnbc_nott.so; do DNBC_B200_LIB=$v python tools/kbench.py scaleout 32000 3; DNBC_B200_LIB=$v python tools/kbench.py largeN 12500 3; done > gpurun_out/k_timing.jsonl 2> gpurun_out/k_err.log
python - <<'PY'
import json
for l in open('gpurun_out/k_timing.jsonl'):
    d=json.loads(l); print(d['lib'], d['workload'], d['ms_per_iter'].get('stats_gmm'), d['ll'])
PY
echo "--- deviation, T-trick"; python tests/param_deviation.py 2>&1 | tail -6
echo "--- deviation, round-2 baseline"; DNBC_B200_LIB=build/libdnbc_nott.so python tests/param_deviation.py 2>&1 | tail -6
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 -k "not two_gpus" > gpurun_out/k_pytest.log 2>&1; tail -5 gpurun_out/k_pytest.log
