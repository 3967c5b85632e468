#!/bin/bash
# ncu --set full on selected kernels at B=64 (K=4): $1 = kernel regex, $2 = skip, $3 = count, $4 = output tag
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$1" -s "$2" -c "$3" -o gpurun_out/prof_$4 -f \
    python tools/ncu_target.py f16x3 64 > gpurun_out/ncu_$4.log 2>&1
tail -n 3 gpurun_out/ncu_$4.log
