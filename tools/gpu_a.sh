#!/bin/bash
# round-2 GPU call A: the whole GPU suite, then the default bench and a short reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt; nproc >> gpurun_out/a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -x --durations=15 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -40 gpurun_out/a_pytest.log
timeout 900 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/a_bench.err
cat gpurun_out/a_bench.json
