#!/bin/bash
mkdir -p gpurun_out
./tools/micro/mma_issue_probe > gpurun_out/mma_issue_probe.jsonl 2>&1; cat gpurun_out/mma_issue_probe.jsonl
for b in 64 8; do
timeout 300 python bench.py --steps 20 --warmup 3 --batch $b --no-cpu-baseline --no-train-leg --no-headline --no-sustained > gpurun_out/y_b$b.json 2> gpurun_out/y_b$b.err
python tools/show_bench.py gpurun_out/y_b$b.json | cut -c1-420; tail -2 gpurun_out/y_b$b.err
done
