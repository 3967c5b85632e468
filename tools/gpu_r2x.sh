#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu > gpurun_out/r2x_pytest_train.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_pytest_train.log; tail -30 gpurun_out/r2x_pytest_train.log
