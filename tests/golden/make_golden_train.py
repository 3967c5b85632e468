"""Golden vectors for the training step from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  The reference's own ``CATRE_disR_shared`` model is
called with ``do_loss=True`` exactly as the training loop does (core/catre/engine/engine.py:293-318): points
re-posed with the current estimate (batching.py:127-140, ZERO_CENTER_INPUT), forward, ``sum(loss_dict.values())``,
``backward()``; the next iteration starts from the detached prediction.  Written to
``tests/golden/golden_train.npz``:

  inputs        the seeded synthetic batch (catre_b200.synth.make_train_batch) + the symmetry rotations the
                reference's data loader attaches to y-symmetric categories (data_loader.py:385-401)
  per iteration every entry of loss_dict, the predicted pose/scale, and a digest of every parameter gradient:
                the whole tensor when it has <= 4096 entries, otherwise (sum, sum|.|, L2 norm) in float64 plus 256
                entries at seeded positions

Usage:  python tests/golden/make_golden_train.py [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402  (shim + config loader + model builder)

N_PTS, BATCH, SEED, N_ITER = 1024, 6, 11, 2
SMALL = 4096
N_SAMPLES = 256


def sample_positions(name: str, numel: int) -> np.ndarray:
    """Seeded positions of the sampled gradient entries of a large tensor (same rule in the test)."""
    seed = int.from_bytes(name.encode()[-8:].rjust(8, b"\0"), "little") % (2 ** 31)
    return np.random.RandomState(seed).randint(0, numel, size=N_SAMPLES)


def grad_digest(name: str, g: torch.Tensor) -> dict:
    g = g.detach().double().flatten()
    if g.numel() <= SMALL:
        return {"full": g.numpy()}
    return {"stats": np.array([g.sum().item(), g.abs().sum().item(), g.norm().item()]),
            "samples": g[torch.from_numpy(sample_positions(name, g.numel()))].numpy()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    mg.install_shim(args.ref)
    from catre_b200 import synth
    from lib.pysixd import misc  # the reference's own symmetry discretisation
    from lib.pysixd.misc import transform_normed_pts_batch

    w = synth.load_weights()
    cfg, model, msg = mg.build_reference_model(args.ref, N_PTS, {k: v.clone() for k, v in w.items()})
    model.train()
    batch, tgt = synth.make_train_batch(BATCH, N_PTS, SEED, round_robin_cls=True)
    sym_rots = np.array([s["R"] for s in misc.get_axis_symmetry_transformations(
        np.array([0, 1, 0]), max_sym_disc_step=cfg.INPUT.MAX_SYM_DISC_STEP)], dtype=np.float32)
    sym_info = [sym_rots if bool(s) else None for s in tgt.sym_y]

    out = {"pcl": batch.pcl.numpy(), "prior_cls": batch.obj_cls.numpy().astype(np.int16),
           "init_pose": batch.init_pose.numpy(), "init_scale": batch.init_scale.numpy(), "K": batch.K.numpy(),
           "gt_pose": tgt.gt_pose.numpy(), "gt_scale": tgt.gt_scale.numpy(), "sym_y": tgt.sym_y.numpy(),
           "sym_rots": sym_rots}
    pose, scale = batch.init_pose, batch.init_scale
    for it in range(1, N_ITER + 1):
        x = (batch.pcl - pose[:, :3, 3].unsqueeze(1)).permute(0, 2, 1)  # batching.py / batch_test.py:95
        tfd = transform_normed_pts_batch(batch.prior, pose[:, :3, :3], t=None, scale=scale).permute(0, 2, 1)
        model.zero_grad(set_to_none=True)
        out_dict, loss_dict = model(x, tfd, init_pose=pose, init_scale=scale, K_zoom=batch.K, obj_class=batch.obj_cls,
                                    gt_ego_rot=tgt.gt_pose[:, :3, :3], gt_trans=tgt.gt_pose[:, :3, 3], gt_scale=tgt.gt_scale,
                                    obj_kps=batch.prior, mean_scales=torch.zeros_like(scale), sym_info=sym_info,
                                    do_loss=True, cur_iter=it)
        total = sum(loss_dict.values())
        total.backward()
        pose, scale = out_dict[f"pose_{it}"].detach(), out_dict[f"scale_{it}"].detach()
        out[f"it{it}_pose"], out[f"it{it}_scale"] = pose.numpy(), scale.numpy()
        out[f"it{it}_loss_names"] = np.array(sorted(loss_dict.keys()))
        out[f"it{it}_loss_values"] = np.array([loss_dict[k].item() for k in sorted(loss_dict.keys())], dtype=np.float64)
        n_none = 0
        for name, p in model.named_parameters():
            if p.grad is None:  # the head's unused norm / conv_p of the config never see a gradient
                n_none += 1
                continue
            for k, v in grad_digest(name, p.grad).items():
                out[f"it{it}_grad/{name}/{k}"] = v
        print(f"iter {it}: total {total.item():.6f}", {k: round(v.item(), 6) for k, v in loss_dict.items()},
              f"params without grad: {n_none}")
    np.savez_compressed(os.path.join(HERE, "golden_train.npz"), **out)
    print("wrote golden_train.npz", msg, f"{os.path.getsize(os.path.join(HERE, 'golden_train.npz')) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
