#!/bin/bash
# round-2 GPU call E10: randomised parity sweep over the session's kernels (points-packed emission, lane-per-sequence
# recursions on every third case, statistics sweep without the software pipeline, static rings of k_fb2)
mkdir -p gpurun_out
FUZZ_LONG=1 timeout 1200 python tests/fuzz_parity.py 700 20261020 > gpurun_out/e10_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -4 gpurun_out/e10_fuzz.log | cut -c1-300
timeout 600 python tests/fuzz_parity.py 12 77 large > gpurun_out/e10_fuzz_large.log 2>&1; echo "fuzz large rc=$?"; tail -2 gpurun_out/e10_fuzz_large.log | cut -c1-300
timeout 600 python tests/fuzz_parity.py 150 99 mle > gpurun_out/e10_fuzz_mle.log 2>&1; echo "fuzz mle rc=$?"; tail -2 gpurun_out/e10_fuzz_mle.log | cut -c1-300
