#!/bin/bash
# GPU visit: ts head as a side lane of the training chain (second stream / graph branch): tests + step time with / without
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_train_gemm_gpu.py tests/test_train_gpu.py -q -m gpu 2>&1 | tail -3
TRAIN_PROBE_MODES=tc,tc-nolanes,tc,tc-nolanes,tc-nograph,tc-nograph-nolanes timeout 300 python tools/train_probe.py 16 64 > gpurun_out/r3g_train_probe_lanes.log 2>&1; cat gpurun_out/r3g_train_probe_lanes.log
