#!/bin/bash
# quick check after a kernel change: stage parity, the parity suite, bench lines at 64 / 256 objects
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_stages_gpu.py -q -s > gpurun_out/pytest_stages.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_stages.log; tail -6 gpurun_out/pytest_stages.log
timeout 900 python -m pytest tests/test_parity_gpu.py -q -s -x > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity.log
grep -E "full-size parity|passed|failed|Error|error" gpurun_out/pytest_parity.log | tail -20
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-train-leg > gpurun_out/bench.json 2> gpurun_out/bench.err; python tools/show_bench.py gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --batch 256 --no-cpu-baseline --no-train-leg > gpurun_out/bench_b256.json 2>> gpurun_out/bench.err; python tools/show_bench.py gpurun_out/bench_b256.json
