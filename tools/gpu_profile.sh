#!/bin/bash
# Round-end evidence: bench lines, ncu launch list of the bench command, ncu --set full of one iteration's big kernels.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --steps 10 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/bench_b256.json 2>> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# second refine call of ncu_target (warm): skip the first call's 4 x 10 matching launches, take one iteration
ncu --set full --clock-control none --import-source on -k regex:'tc_gemm|rot_fused|enc_fused|rot_tail' -s 40 -c 10 -o gpurun_out/prof_iter -f \
    python tools/ncu_target.py f16x3 64 > gpurun_out/ncu_iter.log 2>&1
tail -n 2 gpurun_out/ncu_iter.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -n 3 gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 --batch 256 --n-pts 2048 --n-iter 8 --no-cpu-baseline > gpurun_out/bench_b256_n2048_k8.json 2>> gpurun_out/bench.err; tail -c 400 gpurun_out/bench_b256_n2048_k8.json
