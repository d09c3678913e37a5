#!/bin/bash
# round-2 very last 1-GPU call at HEAD: whole GPU suite + two more long fuzz seeds
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "not two_gpus" > gpurun_out/r02fin4_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02fin4_pytest_gpu.log; tail -3 gpurun_out/r02fin4_pytest_gpu.log
for seed in 31337 424242; do
  FUZZ_LONG=1 timeout 600 python tests/fuzz_parity.py 700 $seed 2>&1 | grep -v "^  File\|^    " | tail -3 > gpurun_out/f4_fuzz_$seed.txt; echo "seed $seed: $(tail -1 gpurun_out/f4_fuzz_$seed.txt)"
done
timeout 300 python tests/fuzz_parity.py 300 777 mle 2>&1 | tail -2
