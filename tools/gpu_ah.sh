#!/bin/bash
# round-2 GPU call AH: max-plus Viterbi tiles for 65..128 states (padded to 256) against the lane-group kernel
mkdir -p gpurun_out
timeout 600 python tests/fuzz_parity.py 200 777 > gpurun_out/ah_fuzz.txt 2>&1; echo "fuzz rc=$?" >> gpurun_out/ah_fuzz.txt; tail -2 gpurun_out/ah_fuzz.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode or viterbi or tensor_core or large_state" > gpurun_out/ah_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ah_pytest.log; tail -2 gpurun_out/ah_pytest.log
rm -f gpurun_out/ah_dbench.jsonl
for w in "custom:N=100,T=200 50000" "custom:N=128,T=200 50000" "custom:N=80,T=200 50000"; do
  for lib in "" build/libdnbc_vtold.so; do
    DNBC_B200_LIB=$lib timeout 300 python tools/dbench.py $w >> gpurun_out/ah_dbench.jsonl 2>> gpurun_out/ah_err.log
  done
done
cut -c1-330 gpurun_out/ah_dbench.jsonl
tail -5 gpurun_out/ah_err.log 2>/dev/null
