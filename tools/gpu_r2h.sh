#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 3; do
for b in 8 64; do
CATRE_FCC_DBG=$dbg timeout 300 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline --no-train-leg --no-headline --no-sustained > gpurun_out/z_d${dbg}_b$b.json 2> gpurun_out/z.err
echo "dbg=$dbg"; python tools/show_bench.py gpurun_out/z_d${dbg}_b$b.json | cut -c1-200; tail -2 gpurun_out/z.err
done
done
