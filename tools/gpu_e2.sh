#!/bin/bash
# round-2 GPU call E12: long randomised parity sweep at HEAD
mkdir -p gpurun_out
FUZZ_LONG=1 timeout 1500 python tests/fuzz_parity.py 2500 7 > gpurun_out/e12_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/e12_fuzz.log | cut -c1-300
timeout 900 python tests/fuzz_parity.py 30 78 large > gpurun_out/e12_fuzz_large.log 2>&1; echo "fuzz large rc=$?"; tail -2 gpurun_out/e12_fuzz_large.log | cut -c1-300
timeout 600 python tests/fuzz_parity.py 400 100 mle > gpurun_out/e12_fuzz_mle.log 2>&1; echo "fuzz mle rc=$?"; tail -2 gpurun_out/e12_fuzz_mle.log | cut -c1-300
