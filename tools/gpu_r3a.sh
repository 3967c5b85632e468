#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/bench_evaluator.py > gpurun_out/r3a_bench_evaluator.log 2>&1; cat gpurun_out/r3a_bench_evaluator.log
