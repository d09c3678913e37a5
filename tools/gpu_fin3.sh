#!/bin/bash
# round-2 last 1-GPU call at HEAD: bench (own arm), then the fuzz sweeps (700 general cases with long sequences,
# 24 at the tensor-core state counts, 150 supervised-learner cases)
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/f3_bench.log 2> gpurun_out/f3_bench.err; echo "bench rc=$?"
grep '^{' gpurun_out/f3_bench.log | tail -1 > gpurun_out/r02fin2_bench_1gpu.json; cut -c1-300 gpurun_out/r02fin2_bench_1gpu.json
FUZZ_LONG=1 timeout 900 python tests/fuzz_parity.py 700 9001 > gpurun_out/r02fin2_fuzz.txt 2>&1; echo "fuzz rc=$?" >> gpurun_out/r02fin2_fuzz.txt; tail -2 gpurun_out/r02fin2_fuzz.txt
timeout 600 python tests/fuzz_parity.py 24 9002 large >> gpurun_out/r02fin2_fuzz.txt 2>&1; echo "fuzz large rc=$?" >> gpurun_out/r02fin2_fuzz.txt; tail -2 gpurun_out/r02fin2_fuzz.txt
timeout 600 python tests/fuzz_parity.py 150 9003 mle >> gpurun_out/r02fin2_fuzz.txt 2>&1; echo "fuzz mle rc=$?" >> gpurun_out/r02fin2_fuzz.txt; tail -2 gpurun_out/r02fin2_fuzz.txt
