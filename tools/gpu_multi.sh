#!/bin/bash
# multi-GPU bench lines: $1 = N GPUs; default workload, then config3 (bf16) and config5 (mixed categories, table entry)
N=$1
mkdir -p gpurun_out
for wl in config2 config3 config5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --workload $wl --no-cpu-baseline --no-train-leg > gpurun_out/bench_${N}gpu_$wl.json 2> gpurun_out/bench_${N}gpu_$wl.err
  python tools/show_bench.py gpurun_out/bench_${N}gpu_$wl.json | head -1; tail -2 gpurun_out/bench_${N}gpu_$wl.err
done
