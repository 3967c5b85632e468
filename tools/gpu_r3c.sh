#!/bin/bash
# GPU visit: source-level stall profile of the training GEMM on its largest forward launch (conv4: M = 32768, N = 1024, K = 512)
# and on a long-K weight-gradient launch
mkdir -p gpurun_out
TRAIN_PROBE_MODES=tc-nograph timeout 400 ncu --set full --import-source on --clock-control none -k regex:"tk_gemm_tc" -s 193 -c 1 -o /tmp/prof_tg_fwd -f python tools/train_probe.py 16 > gpurun_out/r3c_ncu.log 2>&1
ncu -i /tmp/prof_tg_fwd.ncu-rep --page source --csv > /tmp/tg_fwd_src.csv 2>/dev/null
python tools/ncu_src.py /tmp/tg_fwd_src.csv 40 > gpurun_out/r3c_ncu_src_tk_gemm_tc_conv4_fwd.txt 2>&1; head -60 gpurun_out/r3c_ncu_src_tk_gemm_tc_conv4_fwd.txt
python tools/ncu_raw.py /tmp/prof_tg_fwd.ncu-rep > gpurun_out/r3c_ncu_raw_conv4_fwd.txt 2>&1; cat gpurun_out/r3c_ncu_raw_conv4_fwd.txt
ncu -i /tmp/prof_tg_fwd.ncu-rep --page details 2>/dev/null | grep -E "Stall|stall|Warp Cycles Per Issued|Issued Warp|Eligible|No Eligible|Theoretical Occ|Achieved Occ|L1/TEX Hit|L2 Hit|Mem Busy|Max Bandwidth|Mem Pipes" | head -30
