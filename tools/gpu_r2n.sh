#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_stages_gpu.py -q -x -m gpu > gpurun_out/pytest_parity_x.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity_x.log; tail -4 gpurun_out/pytest_parity_x.log
for mode in split fused; do
for b in 64 256 8; do
CATRE_ROT_TAIL=$mode timeout 120 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline --no-train-leg --no-headline --no-sustained > gpurun_out/t_${mode}_b$b.json 2> gpurun_out/z.err
echo "== $mode"; python tools/show_bench.py gpurun_out/t_${mode}_b$b.json | cut -c1-330; tail -2 gpurun_out/z.err
done
CATRE_ROT_TAIL=$mode timeout 200 python bench.py --workload config4 --steps 5 --warmup 3 --no-cpu-baseline --no-train-leg --no-headline --no-sustained > gpurun_out/t_${mode}_c4.json 2> gpurun_out/z.err
python tools/show_bench.py gpurun_out/t_${mode}_c4.json | head -1; tail -2 gpurun_out/z.err
done
