#!/bin/bash
# round-2 GPU call I: GPU suite after the fused E-step dispatch rule + forced-path tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
tail -30 gpurun_out/i_pytest.log
: > gpurun_out/i_configs.jsonl
for w in "readme 1000" "robot 200" "gmm100k 100000"; do timeout 300 python tools/kbench.py $w 5 >> gpurun_out/i_configs.jsonl 2>> gpurun_out/i_err.log; done
cat gpurun_out/i_configs.jsonl
