"""Fixture for the GPU training tests: the training ORACLE (oracle/train_oracle.py, pinned to the reference by
tests/test_train_oracle.py) run in float64 on the seeded batch of golden_train.npz, so the GPU box need not spend
CPU minutes on it.  float64 because the fp32 torch run is itself only good to ~3e-3 on the STN gradients (arg-max
near-ties); the CUDA chain is compared with the fp64 values.  Same digest format as make_golden_train.py.

Usage:  python tests/golden/make_oracle_train_fp64.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from make_golden_train import BATCH, N_PTS, SEED, grad_digest  # noqa: E402
from catre_b200 import synth  # noqa: E402
from oracle import train_oracle as to  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    w = {k: v.double() for k, v in synth.load_weights().items()}
    batch, tgt = synth.make_train_batch(BATCH, N_PTS, SEED, round_robin_cls=True)
    rots = to.y_symmetry_rotations()
    sym_info = [rots.astype(np.float64) if s else None for s in tgt.sym_y]
    args = [t.double() for t in (batch.pcl, batch.prior, batch.init_pose, batch.init_scale, batch.K, tgt.gt_pose, tgt.gt_scale)]
    pose, scale, losses, grads = to.train_step(w, *args, sym_info)
    out = {"pose": pose.numpy(), "scale": scale.numpy(), "loss_names": np.array(sorted(losses)),
           "loss_values": np.array([losses[k] for k in sorted(losses)])}
    for name, g in grads.items():
        for k, v in grad_digest(name, g).items():
            out[f"grad/{name}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "oracle_train_fp64.npz"), **out)
    print("wrote oracle_train_fp64.npz", {k: round(v, 6) for k, v in losses.items()})


if __name__ == "__main__":
    main()
