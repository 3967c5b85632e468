"""Golden vectors for the fused Ranger step: the reference's own optimiser (lib/torch_utils/solver/ranger.py, unmodified,
imported from /root/reference -- it needs only torch) run for 14 steps on small tensors of 1, 2 and 3 dimensions with seeded
gradients (a NaN / inf entry is sanitised first, as the training loop does: core/catre/engine/engine.py:349-352).  14 steps
cross both the RAdam rectification threshold (N_sma > 5 from step 6 with beta2 = 0.999) and two lookahead syncs (k = 6).
Writes tests/golden/golden_ranger.npz.   Usage: python tests/golden/make_golden_ranger.py [--ref /root/reference]"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SHAPES = [(7,), (5, 9), (4, 6, 1), (3, 11), (1,)]
N_STEPS = 14


def make_inputs():
    g = torch.Generator().manual_seed(77)
    params = [torch.randn(s, generator=g) for s in SHAPES]
    grads = [[torch.randn(s, generator=g) * (0.1 + 0.05 * k) for s in SHAPES] for k in range(N_STEPS)]
    grads[3][1][2, 4] = float("nan")
    grads[8][0][5] = float("inf")
    grads[8][2][1, 3, 0] = float("-inf")
    return params, grads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    sys.path.insert(0, os.path.join(args.ref, "lib", "torch_utils", "solver"))
    import ranger  # the reference's file, unmodified

    params, grads = make_inputs()
    ps = [torch.nn.Parameter(p.clone()) for p in params]
    opt = ranger.Ranger([{"params": ps[:3], "lr": 1e-2}, {"params": ps[3:], "lr": 3e-3, "weight_decay": 0.1}], lr=1e-2, weight_decay=0)
    out = {}
    for k in range(N_STEPS):
        for p, g in zip(ps, grads[k]):
            p.grad = g.clone()
            torch.nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)  # engine.py:349-352
        opt.step()
        for i, p in enumerate(ps):
            out[f"step{k + 1}_p{i}"] = p.detach().numpy().copy()
    for i, p in enumerate(ps):
        st = opt.state[p]
        out[f"final_exp_avg{i}"], out[f"final_exp_avg_sq{i}"] = st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
        out[f"final_slow{i}"] = st["slow_buffer"].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "golden_ranger.npz"), **out)
    print("wrote golden_ranger.npz;", "p0 after 14 steps:", ps[0].detach().numpy()[:3])


if __name__ == "__main__":
    main()
