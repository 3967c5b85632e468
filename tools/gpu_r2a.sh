#!/bin/bash
# Round 2, visit A: smoke, per-stage parity, the whole GPU suite (incl. the full-size reference goldens and the
# launch-size independence test), bench lines at 64 / 256 objects, and the round-1 leftovers of the training step.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -8 gpurun_out/smoke.log
timeout 300 python -m pytest tests/test_stages_gpu.py -q -s > gpurun_out/pytest_stages.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_stages.log; tail -12 gpurun_out/pytest_stages.log
timeout 1500 python -m pytest tests -m gpu -q -s --deselect tests/test_stages_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "full-size parity|passed|failed|Error|error" gpurun_out/pytest_gpu.log | tail -40
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-train-leg > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --batch 256 --no-cpu-baseline --no-train-leg > gpurun_out/bench_b256.json 2>> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench_b256.json
timeout 200 python tools/train_probe.py 16 64 > gpurun_out/train_probe.log 2>&1; cat gpurun_out/train_probe.log
timeout 100 python tools/optim_probe.py > gpurun_out/optim_probe.log 2>&1; cat gpurun_out/optim_probe.log
timeout 200 python tools/train_loop_probe.py 16 > gpurun_out/train_loop_probe.log 2>&1; cat gpurun_out/train_loop_probe.log
ls -la gpurun_out
