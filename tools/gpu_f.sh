#!/bin/bash
# round-2 GPU call F: fused small-N E-step -- GPU suite, then per-config timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -40 gpurun_out/f_pytest.log
: > gpurun_out/f_configs.jsonl
for w in "readme 1000" "robot 200" "gmm100k 100000"; do timeout 300 python tools/kbench.py $w 5 >> gpurun_out/f_configs.jsonl 2>> gpurun_out/f_err.log; done
cat gpurun_out/f_configs.jsonl; tail -5 gpurun_out/f_err.log
