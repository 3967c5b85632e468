#!/bin/bash
# GPU visit (2 GPUs): the driver's multi-GPU launch of both bench arms on the final code
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r3d_bench_2gpu.json 2> gpurun_out/r3d_bench_2gpu.err; echo "rc=$?"; tail -c 1500 gpurun_out/r3d_bench_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r3d_bench_2gpu_reference.json 2> gpurun_out/r3d_bench_2gpu_reference.err; echo "rc=$?"; tail -c 600 gpurun_out/r3d_bench_2gpu_reference.json
