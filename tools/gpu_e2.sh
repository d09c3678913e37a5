#!/bin/bash
# round-2 GPU call E9: GPU suite after the removal of the component-packed emission kernel; timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "not two_gpus" > gpurun_out/e9_pytest.log 2>&1; tail -4 gpurun_out/e9_pytest.log
for w in "scaleout 32000 3" "largeN 12500 3" "gmm100k 100000 3" "readme 1000 3" "robot 200 3"; do python tools/kbench.py $w 2>> gpurun_out/e9_err.log | cut -c1-300; done
