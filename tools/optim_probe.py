"""GPU probe of the optimiser step: catre_b200.optim.FusedRanger (two launches over all tensors) beside
a per-tensor torch implementation issuing the reference Ranger's op sequence (lib/torch_utils/solver/ranger.py:118-200:
GC mean/sub, two moment updates, sqrt/add/addcdiv, copy, lookahead every k steps) plus the loop's per-tensor nan_to_num
(core/catre/engine/engine.py:349-352), on the model's 68 trained tensors.  Measurement tool; prints one JSON line each.
Usage (GPU box): python tools/optim_probe.py"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catre_b200 import dropin, optim, synth  # noqa: E402


def torch_ranger_step(ps, state, step, lr=1e-4, betas=(0.95, 0.999), eps=1e-5, alpha=0.5, k=6):
    rect, step_size = optim.radam_step_size(step, betas[0], betas[1], 5)
    for p in ps:
        torch.nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)
        g = p.grad
        m, v, slow = state[p]
        if g.dim() > 1:
            g.add_(-(g.mean(dim=tuple(range(1, g.dim())), keepdim=True)))
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        if rect:
            p.data.addcdiv_(m, v.sqrt().add_(eps), value=-step_size * lr)
        else:
            p.data.add_(m, alpha=-step_size * lr)
        if step % k == 0:
            slow.add_(p.data - slow, alpha=alpha)
            p.data.copy_(slow)


def main():
    dev = "cuda"
    w = synth.load_weights()
    names = [n for n in w if n not in dropin.UNUSED_PARAMS]
    out = {}
    for which in ("torch_per_tensor", "fused"):
        ps = [torch.nn.Parameter(w[n].clone().to(dev)) for n in names]
        g = torch.Generator(device=dev).manual_seed(0)
        grads = [torch.randn(p.shape, device=dev, generator=g) * 1e-3 for p in ps]
        if which == "fused":
            opt = optim.FusedRanger(ps, lr=1e-4, nan_to_num=True)
            step = lambda i: opt.step()
        else:
            state = {p: (torch.zeros_like(p), torch.zeros_like(p), p.detach().clone()) for p in ps}
            step = lambda i: torch_ranger_step(ps, state, i)
        times = []
        for i in range(1, 25):
            for p, gr in zip(ps, grads):
                p.grad = gr.clone()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with torch.no_grad():
                step(i)
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
        times = sorted(times[6:])  # past the rectification switch and the first lookahead sync
        out[which] = times[len(times) // 2]
        print(json.dumps({"probe": "optimizer_step", "impl": which, "tensors": len(ps), "elements": sum(p.numel() for p in ps),
                          "ms_per_step_median": out[which]}), flush=True)
    print(json.dumps({"probe": "optimizer_step", "speedup": out["torch_per_tensor"] / out["fused"]}))


if __name__ == "__main__":
    main()
