#!/bin/bash
# round-2 GPU call R (4 GPUs): the driver's launch line at N = 4, both arms
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r_gpus.txt; nproc >> gpurun_out/r_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r_bench4.log 2> gpurun_out/r_bench4.err; echo "bench rc=$?"
grep '^{' gpurun_out/r_bench4.log | tail -1 | cut -c1-1500
tail -3 gpurun_out/r_bench4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r_ref4.log 2>&1
grep '^{' gpurun_out/r_ref4.log | tail -1 | cut -c1-300
