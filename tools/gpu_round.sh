#!/bin/bash
# One GPU-box visit: parity tests, bench line, reference arm, ncu launch list + full capture of the tc GEMMs.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
python bench.py --steps 10 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/bench_b256.json 2>> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 12 -c 12 -o gpurun_out/prof_tc \
    python tools/ncu_target.py f16x3 64 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
