#!/bin/bash
# GPU visit: experiment -- maximum shared-memory carve-out for every kernel of the training chain (no L1 / shared split switch
# around the 193 KB tensor-core GEMM launches)
mkdir -p gpurun_out
TRAIN_PROBE_MODES=tc,tc-carve,tc,tc-carve timeout 300 python tools/train_probe.py 16 64 > gpurun_out/r2u_train_probe_carve.log 2>&1; cat gpurun_out/r2u_train_probe_carve.log
CATRE_TRAIN_CARVEOUT=1 timeout 300 python -m pytest tests/test_train_gpu.py -q -m gpu -x 2>&1 | tail -3
CATRE_TRAIN_CARVEOUT=1 TRAIN_PROBE_MODES=tc-carve-nograph timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1300 --csv --log-file gpurun_out/r2u_train_launches_warm_carve.csv python tools/train_probe.py 16 > /dev/null 2>&1
python tools/train_launch_summary.py gpurun_out/r2u_train_launches_warm_carve.csv > gpurun_out/r2u_train_launch_summary_warm_carve.txt; head -14 gpurun_out/r2u_train_launch_summary_warm_carve.txt
