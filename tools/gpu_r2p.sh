#!/bin/bash
# GPU visit: whole -m gpu suite (all failures listed), smoke(), default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2p_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest_gpu.log; tail -25 gpurun_out/r2p_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2p_smoke.log; tail -12 gpurun_out/r2p_smoke.log
timeout 600 python bench.py > gpurun_out/r2p_bench_default.json 2> gpurun_out/r2p_bench_default.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2p_bench_default.json
