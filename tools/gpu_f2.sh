#!/bin/bash
# round-2 final 2-GPU call: the 2-GPU tests, then the driver's launch line at N = 2 (both arms)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/f2_gpus.txt
timeout 900 python -m pytest tests -m gpu -q -k "two_gpus or communicator or abi_smoke" > gpurun_out/r02f_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest_2gpu.log
tail -5 gpurun_out/r02f_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/f2_bench2.log 2> gpurun_out/f2_bench2.err; echo "bench rc=$?"
grep '^{' gpurun_out/f2_bench2.log | tail -1 > gpurun_out/r02f_bench_2gpu.json; cut -c1-1200 gpurun_out/r02f_bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/f2_ref2.log 2>&1
grep '^{' gpurun_out/f2_ref2.log | tail -1 > gpurun_out/r02f_bench_reference_2gpu.json; cut -c1-500 gpurun_out/r02f_bench_reference_2gpu.json
