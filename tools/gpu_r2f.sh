#!/bin/bash
# Round-2 experiment visit: rot_fused epilogue schedules (CATRE_ROT_VAR), a1T cache policy (CATRE_A1_POLICY) and object-group
# launches of conv3 -> conv4 (CATRE_TRUNK_GROUP).  Every switch leaves the results bit-identical: the parity subset runs under
# the most aggressive combination, the bench lines tell which one pays.
mkdir -p gpurun_out
run() {  # name, env...
  local name=$1; shift
  for b in 64 256; do
    env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline --no-train-leg --no-headline --no-sustained \
      > gpurun_out/x_${name}_b$b.json 2> gpurun_out/x_${name}_b$b.err
    echo "== $name $*"; python tools/show_bench.py gpurun_out/x_${name}_b$b.json | cut -c1-420; tail -2 gpurun_out/x_${name}_b$b.err
  done
}
CATRE_ROT_VAR=2 CATRE_A1_POLICY=3 CATRE_TRUNK_GROUP=16 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_stages_gpu.py -q -x -m gpu \
  > gpurun_out/pytest_parity_x.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity_x.log; tail -4 gpurun_out/pytest_parity_x.log
run base   CATRE_ROT_VAR=0 CATRE_A1_POLICY=0 CATRE_TRUNK_GROUP=0
run rot1   CATRE_ROT_VAR=1 CATRE_A1_POLICY=0 CATRE_TRUNK_GROUP=0
run rot2   CATRE_ROT_VAR=2 CATRE_A1_POLICY=0 CATRE_TRUNK_GROUP=0
run a1rev  CATRE_ROT_VAR=2 CATRE_A1_POLICY=1 CATRE_TRUNK_GROUP=0
run a1keep CATRE_ROT_VAR=2 CATRE_A1_POLICY=3 CATRE_TRUNK_GROUP=0
run tg16   CATRE_ROT_VAR=2 CATRE_A1_POLICY=0 CATRE_TRUNK_GROUP=16
run tg8    CATRE_ROT_VAR=2 CATRE_A1_POLICY=0 CATRE_TRUNK_GROUP=8
run tg32   CATRE_ROT_VAR=2 CATRE_A1_POLICY=0 CATRE_TRUNK_GROUP=32
run all    CATRE_ROT_VAR=2 CATRE_A1_POLICY=3 CATRE_TRUNK_GROUP=16
