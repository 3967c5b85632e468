#!/bin/bash
mkdir -p gpurun_out
for seed in 21 22 23 24; do timeout 200 python tools/train_case_probe.py 3 384 $seed; done > gpurun_out/r2y_case_seeds_384.log 2>&1; cat gpurun_out/r2y_case_seeds_384.log
for seed in 21 22; do timeout 200 python tools/train_case_probe.py 3 1024 $seed; done > gpurun_out/r2y_case_seeds_1024.log 2>&1; cat gpurun_out/r2y_case_seeds_1024.log
