#!/bin/bash
# ncu --set full of the FC-chain kernel with 16-CTA (default) and 8-CTA clusters
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fc_chain -s 3 -c 3 -o gpurun_out/prof_fcc16 -f python tools/ncu_target.py f16x3 64 > gpurun_out/ncu_fcc16.log 2>&1
CATRE_FC_RANKS=8 ncu --set full --clock-control none --import-source on -k regex:fc_chain -s 3 -c 3 -o gpurun_out/prof_fcc8 -f python tools/ncu_target.py f16x3 64 > gpurun_out/ncu_fcc8.log 2>&1
tail -2 gpurun_out/ncu_fcc16.log gpurun_out/ncu_fcc8.log
