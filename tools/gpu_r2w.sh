#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/refine_graph_probe.py > gpurun_out/r2w_refine_graph_probe.log 2>&1; cat gpurun_out/r2w_refine_graph_probe.log
