#!/bin/bash
# round-2 GPU call AA: tcgen05 recursions at every state count 65..512 (multiple-of-64 padding): parity tests, then
# old (Npad = 256 / 512 only) against new per-kernel timings at in-between state counts
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core or large_state or estep_statistics or dispatch" > gpurun_out/aa_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/aa_pytest.log; tail -5 gpurun_out/aa_pytest.log
rm -f gpurun_out/aa_kbench.jsonl
for w in "custom:N=100,T=200 50000" "custom:N=200,T=200 25000" "custom:N=300,T=200 12500" "custom:N=500,T=200 12500" "largeN 12500"; do
  for lib in "" build/libdnbc_tcold.so; do
    DNBC_B200_LIB=$lib timeout 300 python tools/kbench.py $w 2 >> gpurun_out/aa_kbench.jsonl 2>> gpurun_out/aa_err.log
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/aa_kbench.jsonl"):
    d = json.loads(l); m = d["ms_per_iter"]
    print(d["lib"], d["workload"], d["seqs"], "fwd", m.get("forward"), "bwd", m.get("backward"), "xi", m.get("xi"), "total", m["total"], "ll", d["ll"])
PY
tail -5 gpurun_out/aa_err.log 2>/dev/null
