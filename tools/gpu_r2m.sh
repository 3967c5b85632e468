#!/bin/bash
# Round-2 visit for the rows either side of the path: cloud producer (several images per call), evaluator loop,
# pair metrics, and the evidence the training step still lacked: compute-sanitizer memcheck + racecheck and ncu summaries.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_cloud.py tests/test_evaluator.py tests/test_nocs_eval.py tests/test_nocs_map.py tests/test_metrics.py -m gpu -q -x > gpurun_out/pytest_neighbours.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_neighbours.log; tail -4 gpurun_out/pytest_neighbours.log
timeout 120 python tools/bench_cloud.py > gpurun_out/bench_cloud.log 2>&1; tail -2 gpurun_out/bench_cloud.log
timeout 200 python tools/bench_evaluator.py > gpurun_out/bench_evaluator.log 2>&1; tail -2 gpurun_out/bench_evaluator.log
timeout 120 python tools/bench_metrics.py > gpurun_out/bench_metrics.log 2>&1; tail -2 gpurun_out/bench_metrics.log
# ---- sanitizer: the inference chain (cross-CTA exchange of the fused rot kernel included) and the training chain
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "known_answer or different_observed" \
  > gpurun_out/sanitizer_memcheck_refine.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_memcheck_refine.log; tail -6 gpurun_out/sanitizer_memcheck_refine.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_train_gpu.py -m gpu -q -x -k "oracle" \
  > gpurun_out/sanitizer_memcheck_train.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_memcheck_train.log; tail -6 gpurun_out/sanitizer_memcheck_train.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_train_gpu.py -m gpu -q -x -k "oracle" \
  > gpurun_out/sanitizer_racecheck_train.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_racecheck_train.log; tail -6 gpurun_out/sanitizer_racecheck_train.log
# ---- ncu: per-kernel durations of one training step, of the cloud producer and of the NOCS metric kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/train_launches.csv \
    python tools/train_probe.py 16 > gpurun_out/ncu_train.log 2>&1
python tools/show_launches.py gpurun_out/train_launches.csv 0 | tail -1
timeout 300 ncu --set full --clock-control none -k regex:"tk_gemm|KMaxBwd|KLoss|KGn" -s 200 -c 40 -o /tmp/prof_train -f python tools/train_probe.py 16 > gpurun_out/ncu_train_full.log 2>&1
python tools/ncu_raw.py /tmp/prof_train.ncu-rep > gpurun_out/ncu_train_kernels.txt 2>&1; head -12 gpurun_out/ncu_train_kernels.txt
timeout 200 ncu --set full --clock-control none -k regex:"cloud_|pair_metrics|match_greedy" -c 30 -o /tmp/prof_nb -f python -m pytest tests/test_cloud.py tests/test_nocs_map.py -m gpu -q -x > gpurun_out/ncu_neighbours.log 2>&1
python tools/ncu_raw.py /tmp/prof_nb.ncu-rep > gpurun_out/ncu_neighbour_kernels.txt 2>&1; head -12 gpurun_out/ncu_neighbour_kernels.txt
ls -la gpurun_out | tail -20
