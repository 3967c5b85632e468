#!/bin/bash
mkdir -p gpurun_out
CATRE_FC_TILED_MIN_ROWS=1 timeout 300 python -m pytest tests/test_stages_gpu.py -q -s > gpurun_out/pytest_stages_tiled.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_stages_tiled.log; tail -4 gpurun_out/pytest_stages_tiled.log
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_nocs_map.py tests/test_nocs_eval.py -q -s -x -m gpu > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity.log
grep -E "full-size parity|passed|failed|Error|error" gpurun_out/pytest_parity.log | tail -12
for mr in 256 128 64; do
  echo "== tiled from $mr rows"
  CATRE_FC_TILED_MIN_ROWS=$mr timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-train-leg --no-headline > gpurun_out/bench_t$mr.json 2> gpurun_out/bench.err; python tools/show_bench.py gpurun_out/bench_t$mr.json; tail -3 gpurun_out/bench.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --batch 256 --no-cpu-baseline --no-train-leg --no-headline > gpurun_out/bench_b256.json 2>> gpurun_out/bench.err; python tools/show_bench.py gpurun_out/bench_b256.json
CATRE_FC_TILED_MIN_ROWS=100000 timeout 300 python bench.py --steps 10 --warmup 3 --batch 16 --no-cpu-baseline --no-train-leg --no-headline > gpurun_out/bench_b16_chain.json 2>> gpurun_out/bench.err; python tools/show_bench.py gpurun_out/bench_b16_chain.json
CATRE_FC_TILED_MIN_ROWS=1 timeout 300 python bench.py --steps 10 --warmup 3 --batch 16 --no-cpu-baseline --no-train-leg --no-headline > gpurun_out/bench_b16_tiled.json 2>> gpurun_out/bench.err; python tools/show_bench.py gpurun_out/bench_b16_tiled.json
