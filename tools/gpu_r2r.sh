#!/bin/bash
# GPU visit (1 GPU): training tests after the two-stage symmetric-copy search, full ncu capture of the tensor-core training GEMM
# (one whole step's 60 launches, kernel-by-kernel mode), evaluator loop after the host-side trimming, 1-GPU DDP probe baseline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_evaluator.py tests/test_nocs_eval.py -q -m gpu > gpurun_out/r2r_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_pytest.log; tail -6 gpurun_out/r2r_pytest.log
TRAIN_PROBE_MODES=tc-nograph timeout 400 ncu --set full --clock-control none -k regex:"tk_gemm_tc" -s 180 -c 60 -o /tmp/prof_tgemm -f python tools/train_probe.py 16 > gpurun_out/r2r_ncu_tgemm.log 2>&1
python tools/ncu_raw.py /tmp/prof_tgemm.ncu-rep > gpurun_out/r2r_ncu_train_gemm_tc.txt 2>&1; head -70 gpurun_out/r2r_ncu_train_gemm_tc.txt
timeout 300 python tools/bench_evaluator.py > gpurun_out/r2r_bench_evaluator.log 2>&1; cat gpurun_out/r2r_bench_evaluator.log
timeout 200 python tools/train_ddp_probe.py 16 > gpurun_out/r2r_train_ddp_1gpu.log 2>&1; tail -2 gpurun_out/r2r_train_ddp_1gpu.log
timeout 200 python tools/train_probe.py 16 > gpurun_out/r2r_train_probe.log 2>&1; cat gpurun_out/r2r_train_probe.log
# warm-cache launch list of one training step (ncu flushes the caches before every kernel by default, which inflates the many
# few-microsecond kernels of the chain): --cache-control none
TRAIN_PROBE_MODES=tc-nograph timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1300 --csv --log-file gpurun_out/r2r_train_launches_warm.csv python tools/train_probe.py 16 > /dev/null 2>&1
python tools/train_launch_summary.py gpurun_out/r2r_train_launches_warm.csv > gpurun_out/r2r_train_launch_summary_warm.txt; head -45 gpurun_out/r2r_train_launch_summary_warm.txt
