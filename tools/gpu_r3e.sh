#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "two_devices or known_answer or capturable" > gpurun_out/r3e_pytest_two_devices.log 2>&1; echo "rc=$?" >> gpurun_out/r3e_pytest_two_devices.log; tail -15 gpurun_out/r3e_pytest_two_devices.log
