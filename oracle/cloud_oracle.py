"""CPU oracle for the observed-cloud producer: depth back-projection + mask + ball crop +
random sampling, as the reference's TEST data loader builds ``instances.pcl``
(core/catre/datasets/data_loader.py:773-799 with the shipped config: SAMPLE_DEPTH_FROM_BALL=True,
DEPTH_SAMPLE_BALL_RATIO=0.6, FPS_SAMPLE=False, OCCLUDE_MASK_TEST=False).

TEST INFRASTRUCTURE ONLY (same rule as catre_oracle.py).  Parity status: PINNED -- tests/test_cloud.py checks
this restatement bit-for-bit against tests/golden/golden_cloud.npz, which tests/golden/make_golden_cloud.py
produced by calling the unmodified reference functions (lib/pysixd/misc.py:360-378 backproject_th,
core/utils/cat_data_utils.py:209-226 sample_bp_depth, :283-318 crop_ball_from_pts / random_sample,
:380-400 crop_ball_from_depth_image).  Written against those semantics with plain torch ops; nothing copied.
"""
from __future__ import annotations

from typing import List, Tuple

import torch


def backproject(depth: torch.Tensor, K) -> torch.Tensor:
    """[H,W] depth (m) -> [H,W,3] camera-frame points: ((u-cx) z / fx, (v-cy) z / fy, z), every op in depth's dtype
    in exactly this order: subtract, multiply by depth, divide by the focal length (misc.py:372-378)."""
    h, w = depth.shape
    v = torch.arange(h, dtype=depth.dtype) - K[1, 2]
    u = torch.arange(w, dtype=depth.dtype) - K[0, 2]
    vv, uu = torch.meshgrid(v, u, indexing="ij")
    return torch.stack((uu * depth / K[0, 0], vv * depth / K[1, 1], depth), dim=2)


def ball_radii(pose: torch.Tensor, scale: torch.Tensor, ratio: float) -> List:
    """The up-to-10 radii crop_ball_from_pts tries (cat_data_utils.py:283-291, :386): r0 = max(ratio*|R s|, 0.05),
    then x1.10 per retry.  When the floor wins the radius is a Python float (double arithmetic), otherwise an fp32
    tensor multiplied in place -- both are reproduced because they round differently."""
    radius = ratio * torch.norm(pose[:, :3] @ scale)
    radius = max(radius, 0.05)
    out = []
    for _ in range(10):
        out.append(radius.clone() if isinstance(radius, torch.Tensor) else radius)
        radius *= 1.10
    return out


def ball_indices(pts: torch.Tensor, center: torch.Tensor, radii: List) -> torch.Tensor:
    """Indices (ascending = pixel order) of the valid points inside the first radius that holds >= 10 of them, the
    10th radius otherwise; all points when even that ball is empty (cat_data_utils.py:284-294)."""
    distance = torch.sqrt(((pts - center) ** 2).sum(-1))
    idx = None
    for r in radii:
        idx = torch.where(distance <= r)[0]
        if len(idx) >= 10:
            break
    if len(idx) == 0:
        idx = torch.where(distance <= 1e9)[0]
    return idx


def sample_cloud(depth_bp: torch.Tensor, mask: torch.Tensor, pose: torch.Tensor, scale: torch.Tensor, num_points: int,
                 ratio: float = 0.6) -> Tuple[torch.Tensor, torch.Tensor]:
    """One object: valid = mask & (z > 0) in row-major pixel order (sample_bp_depth), ball crop, indices doubled
    until there are >= num_points, torch.randperm on the GLOBAL CPU generator (random_sample).  Returns
    (pcl [num_points,3], selected indices into the valid-point list before sampling)."""
    valid = torch.logical_and(mask, depth_bp[:, :, -1] > 0).flatten().nonzero().squeeze(1)
    pts = depth_bp.reshape(-1, 3)[valid]
    idx = ball_indices(pts, pose[:, 3], ball_radii(pose, scale, ratio))
    if len(idx) == 0:
        raise ValueError("object has no valid depth pixel under its mask (the reference recurses forever here)")
    sel = idx
    while len(idx) < num_points:
        idx = torch.cat([idx, idx], dim=0)
    pick = torch.randperm(len(idx))[:num_points]
    return pts[idx[pick]], sel


def sample_clouds(depth: torch.Tensor, K, masks: torch.Tensor, poses: torch.Tensor, scales: torch.Tensor, num_points: int,
                  ratio: float = 0.6) -> torch.Tensor:
    """All objects of one image, in instance order (the loop at data_loader.py:778-799) -> [B, num_points, 3] fp32."""
    bp = backproject(depth, K)
    return torch.stack([sample_cloud(bp, m, p, s, num_points, ratio)[0].to(torch.float32) for m, p, s in zip(masks, poses, scales)])
