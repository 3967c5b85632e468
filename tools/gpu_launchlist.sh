#!/bin/bash
# per-launch device time of one warm refine call (B=$1, K=4), L2 state as in a real run (--cache-control none)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 108 -c 120 --csv \
    --log-file gpurun_out/launches_b$1.csv python tools/ncu_target.py f16x3 $1 > gpurun_out/ncu_ll.log 2>&1
tail -n 2 gpurun_out/ncu_ll.log
