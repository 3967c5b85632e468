"""The reference's training step timed beside catre_train_step: the pure-torch restatement of the reference modules
(oracle/train_oracle.py: same Conv1d / Linear / bmm / GroupNorm / GELU ops, autograd backward, shipped losses; measurement
tool only) in eager mode on the GPU, with cuDNN/cuBLAS TF32 as PyTorch defaults it and with TF32 off (the fp32-parity
comparison).  Prints one JSON line per configuration.   Usage: python tools/train_ref_probe.py [--device cuda] [B ...]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catre_b200 import synth  # noqa: E402
from oracle import train_oracle as to  # noqa: E402


def main():
    args = sys.argv[1:]
    device = "cuda"
    if args and args[0] == "--device":
        device, args = args[1], args[2:]
    sizes = [int(a) for a in args] or [16, 64]
    w = {k: v.to(device) for k, v in synth.load_weights().items()}
    rots = to.y_symmetry_rotations()
    for B in sizes:
        batch, tgt = synth.make_train_batch(B, 1024, 3, round_robin_cls=True)
        d = batch.to(device)
        gp, gs = tgt.gt_pose.to(device), tgt.gt_scale.to(device)
        sym_info = [rots if s else None for s in tgt.sym_y]
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            step = lambda: to.train_step(w, d.pcl, d.prior, d.init_pose, d.init_scale, d.K, gp, gs, sym_info)
            n_warm, n = (2, 4) if device == "cuda" else (0, 1)
            for _ in range(n_warm):
                step()
            if device == "cuda":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                step()
            if device == "cuda":
                torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3 / n
            print(json.dumps({"probe": "reference_train_step_torch_eager", "device": device, "B": B, "N": 1024, "tf32": tf32,
                              "ms_per_step": ms, "objects_per_s": B / (ms / 1e3),
                              "note": "wall clock around synchronised steps (the step contains host work: the closest-symmetry "
                                      "search runs on the CPU in the reference too, core/utils/pose_utils.py:499-528)"}), flush=True)


if __name__ == "__main__":
    main()
