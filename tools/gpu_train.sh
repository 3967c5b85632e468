#!/bin/bash
# One GPU-box visit for the training step: parity tests, timings (tiled / v2 / naive GEMM; fused vs
# per-tensor optimiser; the reference's torch step), an ncu launch list and a full capture of one step.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_train.sh'
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_train_gpu.py tests/test_train_optim_gpu.py -q > gpurun_out/pytest_train.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_train.log; tail -15 gpurun_out/pytest_train.log
timeout 200 python tools/train_probe.py 16 64 > gpurun_out/train_probe.log 2>&1; cat gpurun_out/train_probe.log
timeout 100 python tools/optim_probe.py > gpurun_out/optim_probe.log 2>&1; cat gpurun_out/optim_probe.log
timeout 200 python tools/train_loop_probe.py 16 > gpurun_out/train_loop_probe.log 2>&1; cat gpurun_out/train_loop_probe.log
timeout 100 python tools/train_ref_probe.py 16 64 > gpurun_out/train_ref_probe.log 2>&1; cat gpurun_out/train_ref_probe.log
# per-kernel durations of one training step (serialised under ncu: compare shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/train_launches.csv \
    python tools/train_probe.py 16 > gpurun_out/ncu_train.log 2>&1
CATRE_TRAIN_GEMM=v1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:tk_gemm -s 40 -c 6 -o gpurun_out/prof_train_gemm \
    python tools/train_probe.py 16 > gpurun_out/ncu_train_full.log 2>&1
ls -la gpurun_out
