#!/bin/bash
# round-2 closing call: streamed E-step with upload pieces inside compute chunks: parity tests, short bench, then the suite
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "streamed or chunked" 2>&1 | tail -2 > gpurun_out/f6_tests.txt; cat gpurun_out/f6_tests.txt
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu --no-decode --no-extras > gpurun_out/f6_bench.log 2> gpurun_out/f6_bench.err; echo "bench rc=$?"
grep '^{' gpurun_out/f6_bench.log | tail -1 > gpurun_out/r02fin6_bench_1gpu_short.json; python -c "
import json; d=json.load(open('gpurun_out/r02fin6_bench_1gpu_short.json')); print('ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'll', d.get('loglik_after'))"
timeout 110 python -m pytest tests -m gpu -q -x -k "not two_gpus and not long_sequences and not fuzz" > gpurun_out/r02fin6_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02fin6_pytest_gpu.log; tail -2 gpurun_out/r02fin6_pytest_gpu.log
