"""CPU oracle for the CATRE training step: losses and parameter gradients.

TEST INFRASTRUCTURE ONLY (same rules as ``oracle/catre_oracle.py``): imported by ``tests/`` only, never by the
product path.

Parity status: PINNED.  ``tests/golden/make_golden_train.py`` runs the unmodified reference model with
``do_loss=True`` the way its training loop does (core/catre/engine/engine.py:293-318), sums the loss dict, calls
``backward()`` and commits the losses and a digest of all 68 parameter gradients (``tests/golden/golden_train.npz``);
``tests/test_train_oracle.py`` checks this restatement against them.

Two independent routes to the gradients live here:
  * ``train_step``       autograd through the functional forward of ``catre_oracle`` + the loss restatement below;
  * ``manual_backward``  the hand-derived backward, stage by stage, in the decomposition the CUDA backward chain
                         uses (rank-1 rot-tail gradient, GroupNorm backward from two group sums, the layer-0 split,
                         sparse max-pool backward through the arg-max points, per-set transform gradients).  It is
                         checked against ``train_step`` so the algebra of every CUDA stage is verified on the CPU.
Each function cites the reference file:line it follows.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import catre_oracle as co

Weights = Dict[str, torch.Tensor]

# parameters of the shipped config that never receive a gradient: the heads' unused `norm` and, in the ts head
# only, nothing else (core/catre/models/heads/conv_out_per_rot_head.py:96-101, fc_trans_size_head.py:33-36)
UNUSED = ("rot_head.rot_head_x.norm.weight", "rot_head.rot_head_x.norm.bias", "rot_head.rot_head_y.norm.weight",
          "rot_head.rot_head_y.norm.bias", "ts_head.norm.weight", "ts_head.norm.bias")


def y_symmetry_rotations(max_sym_disc_step: float = 0.01) -> np.ndarray:
    """Discretised rotations about y the data loader attaches to a y-symmetric object
    (lib/pysixd/misc.py:220-231 with INPUT.MAX_SYM_DISC_STEP = 0.01, configs/_base_/catre_base.py:24):
    i * 2 pi / n for i = 1 .. n-1, n = ceil(pi / step).  fp32 [n-1, 3, 3] (data_loader.py:397)."""
    n = int(np.ceil(np.pi / max_sym_disc_step))
    out = np.zeros((n - 1, 3, 3), dtype=np.float64)
    for i in range(1, n):
        a = i * 2.0 * np.pi / n
        c, s = np.cos(a), np.sin(a)
        out[i - 1] = [[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]]
    return out.astype(np.float32)


def _rot_error_deg(r_est: np.ndarray, r_gt: np.ndarray) -> float:
    """lib/pysixd/pose_error.py:359-374 (rotation error in degrees, clamped cosine)."""
    tr = np.trace(np.dot(r_est, r_gt.T))
    tr = tr if tr <= 3 else 3
    return float(np.rad2deg(np.arccos(min(1.0, max(-1.0, 0.5 * (tr - 1.0))))))


def closest_sym_rot(pred_rot: torch.Tensor, gt_rot: torch.Tensor, sym_info: List[Optional[np.ndarray]]) -> torch.Tensor:
    """get_closest_rot_batch (core/utils/pose_utils.py:472-528): per object, the ground-truth rotation or the
    symmetric variant R_gt . S_i with the smallest rotation error to the (detached) prediction; strict '<', so the
    plain ground truth wins ties."""
    out = gt_rot.detach().clone().numpy()
    pred = pred_rot.detach().numpy()
    for b, sym in enumerate(sym_info):
        if sym is None:
            continue
        best, best_err = out[b].copy(), _rot_error_deg(pred[b], out[b])
        gt_b = out[b].copy()
        for s in np.asarray(sym).reshape(-1, 3, 3):
            cand = gt_b.dot(s)
            err = _rot_error_deg(pred[b], cand)
            if err < best_err:
                best, best_err = cand, err
        out[b] = best
    return torch.from_numpy(out).to(gt_rot.dtype)


def catre_loss(rot: torch.Tensor, trans: torch.Tensor, scale: torch.Tensor, gt_rot: torch.Tensor, gt_trans: torch.Tensor,
               gt_scale: torch.Tensor, kps: torch.Tensor, sym_info: List[Optional[np.ndarray]]) -> Dict[str, torch.Tensor]:
    """CATRE_disR_shared.catre_loss with the shipped LOSS_CFG (core/catre/models/CATRE_disR_shared.py:168-288;
    configs/catre/NOCS_REAL/aug05_..._120e.py:115-134: symmetric point-matching loss on R only with scale, L1;
    angular rotation loss for asymmetric objects, L1 on the y axis for symmetric ones; L1 on xy / z; L1 on scale;
    every weight 1)."""
    loss: Dict[str, torch.Tensor] = {}
    # point matching (core/catre/losses/pm_loss.py:110-130): R (s * kps) vs R_gt* (s_gt * kps), L1 mean, times 3
    gt_sym = closest_sym_rot(rot, gt_rot, sym_info)
    est = (kps * scale.unsqueeze(1)) @ rot.transpose(1, 2)
    tgt = (kps * gt_scale.unsqueeze(1)) @ gt_sym.transpose(1, 2)
    loss["loss_PM_R"] = 3.0 * (est - tgt).abs().mean()
    # rotation (CATRE_disR_shared.py:222-250; core/catre/losses/rot_loss.py:45-58)
    is_sym = torch.tensor([s is not None for s in sym_info])
    if (~is_sym).any():
        m = rot[~is_sym] @ gt_rot[~is_sym].transpose(1, 2)
        cos = (m.diagonal(dim1=1, dim2=2).sum(1) - 1.0) / 2.0
        loss["loss_rot"] = ((1.0 - cos) / 2.0).mean()
    if is_sym.any():
        loss["loss_yaxis_rot"] = (rot[is_sym][:, :, 1] - gt_rot[is_sym][:, :, 1]).abs().mean()
    # translation, disentangled (CATRE_disR_shared.py:253-262) and scale (:277-286)
    loss["loss_trans_xy"] = (trans[:, :2] - gt_trans[:, :2]).abs().mean()
    loss["loss_trans_z"] = (trans[:, 2] - gt_trans[:, 2]).abs().mean()
    loss["loss_scale"] = (scale - gt_scale).abs().mean()
    return loss


def train_step(w: Weights, pcl: torch.Tensor, kps: torch.Tensor, pose: torch.Tensor, scale: torch.Tensor, K: torch.Tensor,
               gt_pose: torch.Tensor, gt_scale: torch.Tensor, sym_info: List[Optional[np.ndarray]]):
    """One refinement iteration of the training loop (core/catre/engine/engine.py:293-352 without the optimiser):
    re-pose the points with the current (detached) estimate, forward, losses, backward of their sum.
    Returns (pose', scale', loss dict of floats, {name: gradient})."""
    wg = {k: v.clone().requires_grad_(k not in UNUSED) for k, v in w.items()}
    x, tfd = co.update_points(pcl, kps, pose, scale)
    new_pose, new_scale = co.forward_once(wg, x, tfd, pose, scale, K)
    losses = catre_loss(new_pose[:, :3, :3], new_pose[:, :3, 3], new_scale, gt_pose[:, :3, :3], gt_pose[:, :3, 3], gt_scale,
                        kps, sym_info)
    sum(losses.values()).backward()
    grads = {k: v.grad for k, v in wg.items() if v.grad is not None}
    return new_pose.detach(), new_scale.detach(), {k: float(v.detach()) for k, v in losses.items()}, grads
