#!/bin/bash
# GPU visit: final state of the round -- whole suite, smoke, bench line, training iteration probe, memcheck of the training chain
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r3h_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3h_pytest_gpu.log; tail -4 gpurun_out/r3h_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3h_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r3h_smoke.log; tail -2 gpurun_out/r3h_smoke.log
timeout 600 python bench.py > gpurun_out/r3h_bench_default.json 2> gpurun_out/r3h_bench_default.err; echo "bench rc=$?"
timeout 200 python tools/train_loop_probe.py 16 > gpurun_out/r3h_train_loop_probe.log 2>&1; cat gpurun_out/r3h_train_loop_probe.log
timeout 200 python tools/train_probe.py 16 64 > gpurun_out/r3h_train_probe.log 2>&1; cat gpurun_out/r3h_train_probe.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_train_gpu.py -m gpu -q -x -k "matches_oracle or graph_replay or other_sizes or lanes" \
  > gpurun_out/r3h_sanitizer_memcheck_train.log 2>&1; echo "rc=$?" >> gpurun_out/r3h_sanitizer_memcheck_train.log; tail -5 gpurun_out/r3h_sanitizer_memcheck_train.log
