#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -m gpu -k "launch_size or full_size or golden" > gpurun_out/pytest_parity.log 2>&1; tail -3 gpurun_out/pytest_parity.log
for mr in 100000 64 16; do
  echo "== fc3 tiled from $mr rows"
  for b in 8 64 256; do
  CATRE_FC3_TILED_MIN_ROWS=$mr timeout 300 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline --no-train-leg --no-headline > gpurun_out/bench_f$mr_$b.json 2> gpurun_out/bench.err; python tools/show_bench.py gpurun_out/bench_f$mr_$b.json | cut -c1-330; tail -3 gpurun_out/bench.err
  done
done
