#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_stages_gpu.py -q -x -m gpu > gpurun_out/pytest_parity_x.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity_x.log; tail -12 gpurun_out/pytest_parity_x.log
for b in 64 256 8; do
timeout 120 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline --no-train-leg --no-headline --no-sustained > gpurun_out/z_b$b.json 2> gpurun_out/z.err
python tools/show_bench.py gpurun_out/z_b$b.json | cut -c1-400; tail -2 gpurun_out/z.err
done
