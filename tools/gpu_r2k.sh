#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_stages_gpu.py -q -x -m gpu > gpurun_out/pytest_parity_x.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity_x.log; tail -4 gpurun_out/pytest_parity_x.log
CATRE_ROT_TS=1 timeout 120 python bench.py --steps 5 --warmup 3 --batch 64 --no-cpu-baseline --no-train-leg --no-headline --no-sustained 2>&1 | grep ROT_TS | cut -c1-900
for b in 64 256; do
timeout 120 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline --no-train-leg --no-headline --no-sustained > gpurun_out/z_b$b.json 2> gpurun_out/z.err
python tools/show_bench.py gpurun_out/z_b$b.json | cut -c1-400; tail -2 gpurun_out/z.err
done
