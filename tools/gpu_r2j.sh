#!/bin/bash
mkdir -p gpurun_out
for tw in 0 1 3; do
for b in 8 64; do
CATRE_FCC_TWICE=$tw timeout 120 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline --no-train-leg --no-headline --no-sustained > gpurun_out/z_b$b.json 2> gpurun_out/z.err
echo "twice=$tw"; python tools/show_bench.py gpurun_out/z_b$b.json | cut -c1-400; tail -2 gpurun_out/z.err
done
done
