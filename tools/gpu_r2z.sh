#!/bin/bash
# GPU visit: compute-sanitizer on the training chain as it is now (tensor-core GEMM, graph replay, rank / range / gather,
# two-stage symmetric-copy search) + the evaluator loop after the per-launch collation
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_train_gpu.py tests/test_train_gemm_gpu.py -m gpu -q -x -k "matches_oracle or graph_replay or other_sizes or set_weights or tc_gemm or split_k or unaligned or column_max" \
  > gpurun_out/r2z_sanitizer_memcheck_train.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_sanitizer_memcheck_train.log; tail -6 gpurun_out/r2z_sanitizer_memcheck_train.log
timeout 700 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_train_gpu.py -m gpu -q -x -k "matches_oracle" \
  > gpurun_out/r2z_sanitizer_racecheck_train.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_sanitizer_racecheck_train.log; tail -6 gpurun_out/r2z_sanitizer_racecheck_train.log
timeout 300 python tools/bench_evaluator.py > gpurun_out/r2z_bench_evaluator.log 2>&1; cat gpurun_out/r2z_bench_evaluator.log
