#!/bin/bash
# GPU visit: training chain after the latency fixes of the merge / index / gather kernels: tests, step time, warm launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_train_gemm_gpu.py -q -m gpu > gpurun_out/r2s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_pytest.log; tail -6 gpurun_out/r2s_pytest.log
timeout 200 python tools/train_probe.py 16 64 > gpurun_out/r2s_train_probe.log 2>&1; cat gpurun_out/r2s_train_probe.log
TRAIN_PROBE_MODES=tc-nograph timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1300 --csv --log-file gpurun_out/r2s_train_launches_warm.csv python tools/train_probe.py 16 > /dev/null 2>&1
python tools/train_launch_summary.py gpurun_out/r2s_train_launches_warm.csv > gpurun_out/r2s_train_launch_summary_warm.txt; head -45 gpurun_out/r2s_train_launch_summary_warm.txt
timeout 200 python tools/train_loop_probe.py 16 > gpurun_out/r2s_train_loop_probe.log 2>&1; head -3 gpurun_out/r2s_train_loop_probe.log
