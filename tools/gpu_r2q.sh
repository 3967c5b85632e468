#!/bin/bash
# GPU visit: CUDA-graph replay of the training chain + one-launch weight refresh: tests, step time, whole-iteration phases
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_train_optim_gpu.py -q -m gpu > gpurun_out/r2q_pytest_train.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_pytest_train.log; tail -15 gpurun_out/r2q_pytest_train.log
timeout 300 python tools/train_probe.py 16 64 > gpurun_out/r2q_train_probe.log 2>&1; cat gpurun_out/r2q_train_probe.log
timeout 300 python tools/train_loop_probe.py 16 > gpurun_out/r2q_train_loop_probe.log 2>&1; cat gpurun_out/r2q_train_loop_probe.log
