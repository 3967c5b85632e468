"""CPU experiment (not product code): which wide layers tolerate fewer MMA products?

Monkey-patches the oracle's point-wise conv so that chosen layers round their operands the way a tensor-core
mode would (fp32 accumulation is kept), runs the K-loop on seeded inputs and reports the max |delta| on
(R, t, s) against the fp64 oracle.  Usage:

    python tools/precision_probe.py --objects 64 --seeds 11,12 --modes f16:pcl_net.conv4 ...

Mode strings: "<fmt>:<layer>[,<layer>...]" with fmt in {f16, bf16, f16x2w (weights split hi+lo, acts single)}.
All other wide layers are emulated as bf16x3 (the product mode).
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catre_b200 import synth  # noqa: E402
from oracle import catre_oracle as orc  # noqa: E402

WIDE = [
    "pcl_net.stn.conv2", "pcl_net.stn.conv3", "pcl_net.fstn.conv1", "pcl_net.fstn.conv2", "pcl_net.fstn.conv3",
    "pcl_net.conv2", "pcl_net.conv3", "pcl_net.conv4",
    "rot_head.rot_head_x.layers.0", "rot_head.rot_head_y.layers.0",
    "rot_head.rot_head_x.layers.3", "rot_head.rot_head_y.layers.3",
]


def split_bf16(x):
    hi = x.to(torch.bfloat16).float()
    lo = (x - hi).to(torch.bfloat16).float()
    return hi, lo


def make_pw(plan):
    def _pw(w, name, x):
        wt, b = w[name + ".weight"], w[name + ".bias"]
        fmt = plan.get(name)
        if x.dtype != torch.float32 or fmt is None:
            return F.conv1d(x, wt, b)
        if fmt == "bf16x3":
            xh, xl = split_bf16(x)
            wh, wl = split_bf16(wt)
            return F.conv1d(xh, wh, b) + F.conv1d(xh, wl) + F.conv1d(xl, wh)
        if fmt == "f16x3":
            xh = x.half().float(); xl = (x - xh).half().float()
            wh = wt.half().float(); wl = (wt - wh).half().float()
            return F.conv1d(xh, wh, b) + F.conv1d(xh, wl) + F.conv1d(xl, wh)
        if fmt == "f16x3s":  # weights scaled by 2^8 before the split (their fp16 residuals leave the subnormal range)
            xh = x.half().float(); xl = (x - xh).half().float()
            ws = wt * 256.0
            wh = ws.half().float(); wl = (ws - wh).half().float()
            return (F.conv1d(xh, wh) + F.conv1d(xh, wl) + F.conv1d(xl, wh)) * (1.0 / 256.0) + b.reshape(1, -1, 1)
        if fmt == "f16":
            return F.conv1d(x.half().float(), wt.half().float(), b)
        if fmt == "bf16":
            return F.conv1d(x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), b)
        if fmt == "f16x2w":
            wh = wt.half().float()
            wl = (wt - wh).half().float()
            xh = x.half().float()
            return F.conv1d(xh, wh, b) + F.conv1d(xh, wl)
        if fmt == "f16x2a":
            xh = x.half().float()
            xl = (x - xh).half().float()
            wh = wt.half().float()
            return F.conv1d(xh, wh, b) + F.conv1d(xl, wh)
        raise ValueError(fmt)
    return _pw


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--objects", type=int, default=32)
    ap.add_argument("--seeds", default="101")
    ap.add_argument("--iters", type=int, default=4)
    ap.add_argument("--modes", nargs="*", default=[])
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    w32 = synth.load_weights()
    w64 = orc.cast_weights(w32, torch.float64)
    orig = orc._pw
    plans = {"bf16x3-all": {}}
    for m in a.modes:
        fmt, layers = m.split(":")
        plans[m] = {}
        for l in layers.split(","):
            hits = [n for n in WIDE if l in n]
            assert hits, l
            for n in hits:
                plans[m][n] = fmt
    for seed in [int(s) for s in a.seeds.split(",")]:
        bt = synth.make_batch(a.objects, 1024, seed)
        args64 = [t.double() for t in (bt.pcl, bt.prior, bt.init_pose, bt.init_scale, bt.K)]
        orc._pw = orig
        p64, s64 = orc.refine(w64, *args64, a.iters)
        for name, over in plans.items():
            plan = {n: "bf16x3" for n in WIDE}
            plan.update(over)
            orc._pw = make_pw(plan)
            p, s = orc.refine(w32, bt.pcl, bt.prior, bt.init_pose, bt.init_scale, bt.K, a.iters)
            e_r = (p[..., :3].double() - p64[..., :3]).abs().max().item()
            e_t = (p[..., 3].double() - p64[..., 3]).abs().max().item()
            e_s = (s.double() - s64).abs().max().item()
            print(f"seed {seed} {name:60s} dR {e_r:.2e} dt {e_t:.2e} ds {e_s:.2e}", flush=True)
    orc._pw = orig


if __name__ == "__main__":
    main()
