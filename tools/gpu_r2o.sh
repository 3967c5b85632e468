#!/bin/bash
# GPU visit: tensor-core training GEMM -- unit test, training-step parity, step time, launch list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_train_gemm_gpu.py -q -x -m gpu > gpurun_out/pytest_gemm_tc.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gemm_tc.log; tail -15 gpurun_out/pytest_gemm_tc.log
timeout 400 python -m pytest tests/test_train_gpu.py tests/test_train_optim_gpu.py -q -x -m gpu > gpurun_out/pytest_train_tc.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_train_tc.log; tail -15 gpurun_out/pytest_train_tc.log
cat gpurun_out/train_grad_errors.json 2>/dev/null | head -20
timeout 200 python tools/train_probe.py 16 64 > gpurun_out/train_probe_tc.log 2>&1; cat gpurun_out/train_probe_tc.log
TRAIN_PROBE_MODES=tc timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/train_launches_tc.csv python tools/train_probe.py 16 > /dev/null 2>&1
python tools/train_launch_summary.py gpurun_out/train_launches_tc.csv | head -50
