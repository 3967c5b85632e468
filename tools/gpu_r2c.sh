#!/bin/bash
# whole GPU suite + the default bench line (all legs) + the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "full-size parity|stage parity|passed|failed|Error|error" gpurun_out/pytest_gpu.log | tail -30
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 1200 gpurun_out/bench_ref.json
