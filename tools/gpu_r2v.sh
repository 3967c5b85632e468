#!/bin/bash
# GPU visit (2 GPUs): data-parallel training iteration under DistributedDataParallel over NCCL, beside the 1-GPU iteration
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_train_gpu.py -q -m gpu -x 2>&1 | tail -2
timeout 200 python tools/train_ddp_probe.py 16 > gpurun_out/r2v_train_ddp_1gpu.log 2>&1; tail -1 gpurun_out/r2v_train_ddp_1gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/train_ddp_probe.py 16 > gpurun_out/r2v_train_ddp_2gpu.log 2>&1; tail -3 gpurun_out/r2v_train_ddp_2gpu.log
