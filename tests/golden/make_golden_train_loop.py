"""Golden vectors for the whole training iteration (forward + losses + backward + NaN guard + Ranger step), three times in a
row, from the UNMODIFIED reference: its model (core/catre/models/CATRE_disR_shared.py, do_loss=True) and its optimiser
(lib/torch_utils/solver/ranger.py) wired as its training loop wires them (core/catre/engine/engine.py:293-352, the three
parameter groups of CATRE_disR_shared.py:292-315).  N = 128 points per set keeps the CPU emulation of the CUDA chain, which
the test runs against these vectors, to seconds.  Writes tests/golden/golden_train_loop.npz.

Usage:  python tests/golden/make_golden_train_loop.py [--ref /root/reference]"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402
from make_golden_train import grad_digest  # noqa: E402  (same digest format, applied to the weights here)

N_PTS, BATCH, SEED, N_STEPS, LR = 128, 3, 31, 3, 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    mg.install_shim(args.ref)
    from catre_b200 import synth
    from lib.pysixd import misc
    from lib.pysixd.misc import transform_normed_pts_batch
    from lib.torch_utils.solver.ranger import Ranger  # the reference's optimiser, unmodified
    from oracle import catre_oracle

    w = catre_oracle.resize_conv_p(synth.load_weights(), N_PTS)
    cfg, model, msg = mg.build_reference_model(args.ref, N_PTS, {k: v.clone() for k, v in w.items()})
    model.train()
    groups = [{"params": [p for p in part.parameters() if p.requires_grad], "lr": LR}
              for part in (model.pcl_net, model.rot_head, model.ts_head)]  # CATRE_disR_shared.py:292-315
    opt = Ranger(groups, lr=LR, weight_decay=0)
    batch, tgt = synth.make_train_batch(BATCH, N_PTS, SEED, round_robin_cls=True)
    sym_rots = np.array([s["R"] for s in misc.get_axis_symmetry_transformations(
        np.array([0, 1, 0]), max_sym_disc_step=cfg.INPUT.MAX_SYM_DISC_STEP)], dtype=np.float32)
    sym_info = [sym_rots if bool(s) else None for s in tgt.sym_y]
    pose, scale = batch.init_pose, batch.init_scale
    x = (batch.pcl - pose[:, :3, 3].unsqueeze(1)).permute(0, 2, 1)
    tfd = transform_normed_pts_batch(batch.prior, pose[:, :3, :3], t=None, scale=scale).permute(0, 2, 1)
    out = {}
    for it in range(N_STEPS):
        out_dict, loss_dict = model(x, tfd, init_pose=pose, init_scale=scale, K_zoom=batch.K, obj_class=batch.obj_cls,
                                    gt_ego_rot=tgt.gt_pose[:, :3, :3], gt_trans=tgt.gt_pose[:, :3, 3], gt_scale=tgt.gt_scale,
                                    obj_kps=batch.prior, mean_scales=torch.zeros_like(scale), sym_info=sym_info, do_loss=True,
                                    cur_iter=1)
        losses = sum(loss_dict.values())
        losses.backward()
        for param in model.parameters():  # engine.py:349-352
            if param.grad is not None:
                torch.nan_to_num(param.grad, nan=0, posinf=1e5, neginf=-1e5, out=param.grad)
        opt.step()
        opt.zero_grad(set_to_none=True)
        out[f"loss{it + 1}"] = np.float64(losses.item())
        print(f"iteration {it + 1}: total loss {losses.item():.6f}")
    for name, p in model.named_parameters():
        for k, v in grad_digest(name, p.detach()).items():
            out[f"weight/{name}/{k}"] = v
        out[f"delta/{name}"] = np.float64((p.detach() - w[name]).abs().max().item())
    np.savez_compressed(os.path.join(HERE, "golden_train_loop.npz"), **out)
    print("wrote golden_train_loop.npz", msg, "largest weight change", max(float(out[k]) for k in out if k.startswith("delta/")))


if __name__ == "__main__":
    main()
