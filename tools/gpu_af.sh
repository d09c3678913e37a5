#!/bin/bash
# round-2 GPU call AF: single-pass backward epilogue of k_fb_tc (normaliser = last step's posterior sum) against the
# two-pass one: parity tests, then same-call timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core" > gpurun_out/af_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/af_pytest.log; tail -3 gpurun_out/af_pytest.log
echo skip-fuzz > gpurun_out/af_fuzz.txt
rm -f gpurun_out/af_kbench.jsonl
for w in "largeN 100000" "custom:N=200,T=200 100000"; do
  for lib in "" build/libdnbc_nopfe.so; do
    DNBC_B200_LIB=$lib timeout 300 python tools/kbench.py $w 2 >> gpurun_out/af_kbench.jsonl 2>> gpurun_out/af_err.log
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/af_kbench.jsonl"):
    d = json.loads(l); m = d["ms_per_iter"]
    print(d["lib"], d["workload"], d["seqs"], "fwd", m.get("forward"), "bwd", m.get("backward"), "xi", m.get("xi"), "total", m["total"], "ll", d["ll"])
PY
tail -5 gpurun_out/af_err.log 2>/dev/null
