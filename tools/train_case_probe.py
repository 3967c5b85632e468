"""Diagnostic: one training step at (B, N) without symmetric objects, tensor-core vs CUDA-core GEMM, gradients against the fp64
oracle and against each other.  Usage (GPU box): python tools/train_case_probe.py B N [seed]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catre_b200 import engine, synth  # noqa: E402
from oracle import train_oracle as to  # noqa: E402  (checker)
from tests.test_train_gpu import y_symmetry_rotations  # noqa: E402


def main():
    B, N = int(sys.argv[1]), int(sys.argv[2])
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 21
    w32 = {k: (v[:, : 2 * N].contiguous() if k.endswith("conv_p.weight") else v) for k, v in synth.load_weights().items()}
    batch, tgt = synth.make_train_batch(B, N, seed, round_robin_cls=True)
    is_sym = np.zeros(B, bool)
    rots = y_symmetry_rotations()
    args64 = [t.double() for t in (batch.pcl, batch.prior, batch.init_pose, batch.init_scale, batch.K, tgt.gt_pose, tgt.gt_scale)]
    _, _, l_ref, g_ref = to.train_step({k: v.double() for k, v in w32.items()}, *args64, [None] * B)
    d = batch.to("cuda")
    x_pm = (d.pcl - d.init_pose[:, :, 3].unsqueeze(1)).contiguous()
    tfd_pm = ((d.prior * d.init_scale.unsqueeze(1)) @ d.init_pose[:, :, :3].transpose(1, 2)).contiguous()
    flats = {}
    for mode in ("tc", "simt"):
        os.environ["CATRE_TRAIN_GEMM"] = mode
        os.environ["CATRE_TRAIN_GRAPH"] = "0"
        eng = engine.Engine(N, 8, "fp32", 0)
        eng.load_weights(w32)
        eng.train_step(x_pm, tfd_pm, d.prior, d.init_pose, d.init_scale, d.K, tgt.gt_pose.cuda(), tgt.gt_scale.cuda(), is_sym, rots)
        flats[mode] = eng.train_grads_flat(1.0).double().cpu()
        offsets, _ = eng.train_grad_layout()
        eng.close()
    rows = []
    for name, t in w32.items():
        if name not in g_ref:
            continue
        want = g_ref[name].flatten()
        sl = slice(offsets[name], offsets[name] + t.numel())
        sc = max(want.abs().max().item(), 1e-12)
        e_tc, e_simt = (flats["tc"][sl] - want).abs(), (flats["simt"][sl] - want).abs()
        rows.append((e_tc.max().item() / sc, e_simt.max().item() / sc, int((e_tc > 1e-3 * sc).sum()), t.numel(), name))
    print(f"B={B} N={N} seed={seed}")
    for r in sorted(rows, reverse=True)[:5]:
        print("tc %.3e  simt %.3e  entries>1e-3: %d of %d  %s" % r)
    # which conv3 channels carry the tensor-core run's large errors (one ReLU / arg-max flip = one channel)
    name = "pcl_net.conv3.weight"
    sl = slice(offsets[name], offsets[name] + w32[name].numel())
    want = g_ref[name].flatten()
    e = (flats["tc"][sl] - want).abs().view(512, 128)
    bad = (e > 1e-3 * want.abs().max()).any(dim=1).nonzero().flatten().tolist()
    print("conv3 channels with tensor-core errors > 1e-3 of max:", bad)


if __name__ == "__main__":
    main()
