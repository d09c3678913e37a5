#!/bin/bash
# round-2 GPU call P: pruned max-plus Viterbi -- parity tests that exercise it, then timing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 -k "large or n512 or config5 or randomised or long_seq or decode or state_count or empty" > gpurun_out/p_pytest.log 2>&1; tail -6 gpurun_out/p_pytest.log
python tools/dbench.py largeN 125000 > gpurun_out/p_decode.jsonl 2>> gpurun_out/p_err.log; cat gpurun_out/p_decode.jsonl
python tools/dbench.py scaleout 32000 >> gpurun_out/p_decode.jsonl 2>> gpurun_out/p_err.log; tail -1 gpurun_out/p_decode.jsonl
