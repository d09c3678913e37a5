#!/bin/bash
# round-2 GPU call AE: whole GPU suite after the any-state-count work (tcgen05 recursions 65..512, xi partial tiles,
# k_stats2 partial slice, round-1 recursion kernels removed) + the fuzz sweep over the extended state-count lists
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "not two_gpus" > gpurun_out/r02ae_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ae_pytest_gpu.log; tail -3 gpurun_out/r02ae_pytest_gpu.log
timeout 900 python tests/fuzz_parity.py 300 4242 > gpurun_out/r02ae_fuzz.txt 2>&1; echo "fuzz rc=$?" >> gpurun_out/r02ae_fuzz.txt; tail -2 gpurun_out/r02ae_fuzz.txt
timeout 600 python tests/fuzz_parity.py 12 4243 large >> gpurun_out/r02ae_fuzz.txt 2>&1; echo "fuzz large rc=$?" >> gpurun_out/r02ae_fuzz.txt; tail -2 gpurun_out/r02ae_fuzz.txt
