#!/bin/bash
# GPU visit: training chain after the second round of latency fixes (rank/range index, pipelined gather, hoisted conv_p sum,
# 4-way fp64 merges, 4-wide GroupNorm kernels): step time, warm launch list, whole -m gpu suite, smoke, default bench line
mkdir -p gpurun_out
timeout 200 python tools/train_probe.py 16 64 > gpurun_out/r2t_train_probe.log 2>&1; cat gpurun_out/r2t_train_probe.log
TRAIN_PROBE_MODES=tc-nograph timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1300 --csv --log-file gpurun_out/r2t_train_launches_warm.csv python tools/train_probe.py 16 > /dev/null 2>&1
python tools/train_launch_summary.py gpurun_out/r2t_train_launches_warm.csv > gpurun_out/r2t_train_launch_summary_warm.txt; head -24 gpurun_out/r2t_train_launch_summary_warm.txt
timeout 200 python tools/train_loop_probe.py 16 > gpurun_out/r2t_train_loop_probe.log 2>&1; cat gpurun_out/r2t_train_loop_probe.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2t_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest_gpu.log; tail -4 gpurun_out/r2t_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2t_smoke.log; tail -3 gpurun_out/r2t_smoke.log
timeout 600 python bench.py > gpurun_out/r2t_bench_default.json 2> gpurun_out/r2t_bench_default.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2t_bench_default.json
