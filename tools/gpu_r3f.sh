#!/bin/bash
# GPU visit: bias gradients folded into the split-K weight-gradient GEMMs: tests + step time with / without
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_train_gemm_gpu.py tests/test_train_gpu.py -q -m gpu 2>&1 | tail -3
TRAIN_PROBE_MODES=tc,tc-nofold,tc,tc-nofold timeout 300 python tools/train_probe.py 16 64 > gpurun_out/r3f_train_probe_fold.log 2>&1; cat gpurun_out/r3f_train_probe_fold.log
