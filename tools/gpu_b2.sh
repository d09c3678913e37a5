#!/bin/bash
# round-2 GPU call B2: full bench line at HEAD (1 GPU) + GPU suite
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/b2_bench.log 2> gpurun_out/b2_bench.err; echo "bench rc=$?"
grep '^{' gpurun_out/b2_bench.log | tail -1 > gpurun_out/r02u_bench_1gpu.json; cut -c1-3000 gpurun_out/r02u_bench_1gpu.json
tail -3 gpurun_out/b2_bench.err
