#!/bin/bash
# parity tests + B=64 / B=256 bench lines (no profiler)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/bench_b256.json 2>> gpurun_out/bench.err; tail -c 300 gpurun_out/bench_b256.json
