#!/bin/bash
# round-2 final 8-GPU call: the driver's launch line at N = 8 (our arm; the CPU arm is the same as at N = 2)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/f8_gpus.txt; nproc >> gpurun_out/f8_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/f8_bench8.log 2> gpurun_out/f8_bench8.err; echo "bench rc=$?"
grep '^{' gpurun_out/f8_bench8.log | tail -1 > gpurun_out/r02f_bench_8gpu.json; cut -c1-1500 gpurun_out/r02f_bench_8gpu.json
tail -3 gpurun_out/f8_bench8.err
