#!/bin/bash
# round-2 GPU call E (2 GPUs): the collective behind the C ABI -- 2-GPU tests, then the 2-rank bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/e_gpus.txt
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q -k "two_gpus or communicator or abi_smoke" > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -15 gpurun_out/e_pytest.log
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/e_bench2.log 2> gpurun_out/e_bench2.err; echo "bench rc=$?"
grep -c "NCCL INFO" gpurun_out/e_bench2.log gpurun_out/e_bench2.err
grep -h "nranks\|Init COMPLETE" gpurun_out/e_bench2.log gpurun_out/e_bench2.err | head -8
grep '^{' gpurun_out/e_bench2.log | tail -1
OMP_NUM_THREADS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/e_ref2.log 2>&1
grep '^{' gpurun_out/e_ref2.log | tail -1 | cut -c1-600
