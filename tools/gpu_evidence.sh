#!/bin/bash
# Final evidence visit (1 GPU): launch list of the bench command, full ncu captures of one iteration at 64 and 256 objects,
# bench lines for the named configs, neighbour benchmarks.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-leg --no-headline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -s 21 -c 21 -o gpurun_out/prof_iter64 -f python tools/ncu_target.py f16x3 64 > gpurun_out/ncu_iter64.log 2>&1
ncu --set full --clock-control none --import-source on -s 21 -c 21 -o gpurun_out/prof_iter256 -f python tools/ncu_target.py f16x3 256 > gpurun_out/ncu_iter256.log 2>&1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --workload config4 --steps 5 --warmup 3 --no-cpu-baseline --no-train-leg > gpurun_out/bench_config4.json 2> gpurun_out/bench_config4.err
python bench.py --batch 256 --steps 10 --warmup 3 --no-cpu-baseline --no-train-leg --no-headline > gpurun_out/bench_b256.json 2> gpurun_out/bench_b256.err
python bench.py --batch 8 --steps 20 --warmup 3 --no-cpu-baseline --no-train-leg --no-headline > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err
python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-train-leg --no-headline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python tools/bench_evaluator.py > gpurun_out/bench_evaluator.log 2>&1
python tools/bench_cloud.py > gpurun_out/bench_cloud.log 2>&1
python tools/bench_metrics.py > gpurun_out/bench_metrics.log 2>&1
ls -la gpurun_out | tail -30
