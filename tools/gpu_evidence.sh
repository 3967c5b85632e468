#!/bin/bash
# Final evidence visit (1 GPU): launch list of the bench command, full ncu captures of one iteration at 64 and 256 objects,
# bench lines for the named configs, neighbour benchmarks.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-leg --no-headline > gpurun_out/ncu_bench.log 2>&1
# the reports are tens of MB each (gpurun_out/ is capped at 64 MiB): summarise them on the box, keep the text
for b in 64 256; do
  ncu --set full --clock-control none -s 17 -c 17 -o /tmp/prof_iter$b -f python tools/ncu_target.py f16x3 $b > gpurun_out/ncu_iter$b.log 2>&1
  python tools/ncu_iter.py /tmp/prof_iter$b.ncu-rep $b --update-traffic > gpurun_out/ncu_iteration_b$b.txt 2>&1
done
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic.json
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
