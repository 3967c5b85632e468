#!/bin/bash
# Round-end evidence: bench lines, ncu launch list of the bench command, ncu --set full of one iteration's big kernels.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --steps 10 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/bench_b256.json 2>> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# second refine call of ncu_target (warm): 27 launches per iteration -> skip the first call (108) and take one iteration
ncu --set full --clock-control none --import-source on -k regex:'tc_gemm|rot_fused|rot_tail' -s 48 -c 12 -o gpurun_out/prof_iter -f \
    python tools/ncu_target.py bf16x3 64 > gpurun_out/ncu_iter.log 2>&1
tail -n 2 gpurun_out/ncu_iter.log
