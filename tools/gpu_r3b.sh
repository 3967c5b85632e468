#!/bin/bash
# GPU visit: tk_gemm_tc with the first loads issued before the TMEM / barrier set-up: unit tests + step time
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_train_gemm_gpu.py tests/test_train_gpu.py -q -m gpu 2>&1 | tail -3
TRAIN_PROBE_MODES=tc,tc-nograph timeout 200 python tools/train_probe.py 16 64 > gpurun_out/r3b_train_probe.log 2>&1; cat gpurun_out/r3b_train_probe.log
