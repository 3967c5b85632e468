#!/bin/bash
# ncu --set full on one iteration's tensor-core launches (+ rot_tail, ts_pose, pw_gemm) at B=64
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 13 -c 13 -o gpurun_out/prof_tc -f \
    python tools/ncu_target.py bf16x3 64 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'rot_tail|ts_pose|pw_gemm|sum_parts|front3' -s 14 -c 14 -o gpurun_out/prof_simt -f \
    python tools/ncu_target.py bf16x3 64 > gpurun_out/ncu_simt.log 2>&1
tail -3 gpurun_out/ncu_full.log gpurun_out/ncu_simt.log
