"""Does replaying the K-loop as a CUDA graph pay at small batches?  Times engine.refine launched kernel by kernel (programmatic
dependent launch between the kernels, the product's path) against a torch.cuda.CUDAGraph replay of the same call, device time
per refine over back-to-back calls.  Usage (GPU box): python tools/refine_graph_probe.py [B ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catre_b200 import engine, synth  # noqa: E402


def timed(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1, 4, 8, 16, 64]
    w = synth.load_weights()
    K = 4
    for B in sizes:
        eng = engine.Engine(1024, max(B, 8), "f16x3", 0)
        eng.load_weights(w)
        b = synth.make_batch(B, 1024, seed=41).to("cuda")
        out = (torch.empty((K + 1, B, 3, 4), device="cuda"), torch.empty((K + 1, B, 3), device="cuda"))
        call = lambda: eng.refine(b.pcl, b.prior, b.init_pose, b.init_scale, b.K, K, out=out)
        call()
        torch.cuda.synchronize()
        ref = (out[0].clone(), out[1].clone())
        eager = timed(call, 50)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                call()
        torch.cuda.current_stream().wait_stream(side)
        replay = timed(g.replay, 50)
        same = bool(torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1]))
        print(json.dumps({"probe": "refine_graph", "B": B, "N": 1024, "K": K, "launches": eng.last_launch_count(),
                          "ms_kernel_by_kernel": eager, "ms_graph_replay": replay, "speedup": eager / replay, "bit_identical": same}), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
