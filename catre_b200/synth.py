"""Seeded synthetic inputs for the pose-refinement path.

The same bytes go to the oracle, the golden generator and the CUDA engine.  Inputs are built from
two small fixtures shipped with the reference and committed under ``tests/golden/``
(``nocs_fixtures.npz``): the six category prior shapes
(reference: datasets/NOCS/obj_models/cr_normed_mean_model_points_spd.pkl) and the 15,374 real
initial (R, t, s) estimates of REAL275
(reference: datasets/NOCS/test_init_poses/init_pose_spd_nocs_real.json).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
FIXTURES_NPZ = os.path.join(GOLDEN_DIR, "nocs_fixtures.npz")
WEIGHTS_NPZ = os.path.join(GOLDEN_DIR, "catre_weights_82cf930e.npz")

# reference: ref/nocs.py:33 (category order) and ref/nocs.py:103 (REAL275 intrinsics)
CATEGORIES = ("bottle", "bowl", "camera", "can", "laptop", "mug")
NOCS_REAL_K = np.array([[591.0125, 0.0, 322.525], [0.0, 590.16775, 244.11084], [0.0, 0.0, 1.0]], dtype=np.float32)


@dataclass
class Fixtures:
    priors: torch.Tensor  # [6, 1024, 3] float64, category order = CATEGORIES
    init_pose: torch.Tensor  # [15374, 3, 4] float64
    init_scale: torch.Tensor  # [15374, 3] float64
    obj_cls: torch.Tensor  # [15374] int64, 0-based


_FIX: Optional[Fixtures] = None


def load_fixtures(path: str = FIXTURES_NPZ) -> Fixtures:
    global _FIX
    if _FIX is None or path != FIXTURES_NPZ:
        z = np.load(path)
        fx = Fixtures(
            priors=torch.from_numpy(z["priors"]),
            init_pose=torch.from_numpy(z["init_pose"]),
            init_scale=torch.from_numpy(z["init_scale"]),
            obj_cls=torch.from_numpy(z["obj_cls"].astype(np.int64)),
        )
        if path != FIXTURES_NPZ:
            return fx
        _FIX = fx
    return _FIX


def load_weights(path: str = WEIGHTS_NPZ) -> Dict[str, torch.Tensor]:
    """The reference's shipped checkpoint (74 fp32 tensors, checkpoint names) as a flat dict."""
    z = np.load(path)
    return {k: torch.from_numpy(z[k]) for k in z.files}


def resize_conv_p(w: Dict[str, torch.Tensor], n_pts: int, n_prior: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Fixture-defined weights for N != 1024 (SURVEY.md 8(d) "Weights per config"): the checkpoint with the
    two conv_p.weight [1, 2*1024, 1] re-sized -- obs half and prior half each linearly re-sampled to n_pts
    and scaled by 1024 / n_pts (conv_p is tied to the point count,
    core/catre/models/heads/conv_out_per_rot_head.py:112)."""
    import torch.nn.functional as F

    out = dict(w)
    for head in ("rot_head.rot_head_x", "rot_head.rot_head_y"):
        cp = w[head + ".conv_p.weight"]
        half = cp.shape[1] // 2
        n_p = n_pts if n_prior is None else n_prior  # the prior half may have its own point count (NUM_KPS != NUM_PCL)
        if half == n_pts and half == n_p:
            continue
        parts = []
        for seg, n in ((cp[:, :half, 0], n_pts), (cp[:, half:, 0], n_p)):
            r = F.interpolate(seg.reshape(1, 1, half).double(), size=n, mode="linear", align_corners=True)
            parts.append(r.reshape(1, n) * (float(half) / float(n)))
        out[head + ".conv_p.weight"] = torch.cat(parts, dim=1).reshape(1, n_pts + n_p, 1).to(cp.dtype).contiguous()
    return out


def resample_prior(prior: torch.Tensor, n_pts: int) -> torch.Tensor:
    """[..., 1024, 3] -> [..., n_pts, 3]: first-N when shrinking, tiling when growing."""
    n0 = prior.shape[-2]
    if n_pts <= n0:
        return prior[..., :n_pts, :].contiguous()
    reps = (n_pts + n0 - 1) // n0
    return torch.cat([prior] * reps, dim=-2)[..., :n_pts, :].contiguous()


def _axis_angle_to_mat(axis: torch.Tensor, angle: torch.Tensor) -> torch.Tensor:
    """Rodrigues; axis [B,3] unit, angle [B] radians -> [B,3,3] (float64)."""
    b = axis.shape[0]
    kx, ky, kz = axis[:, 0], axis[:, 1], axis[:, 2]
    zero = torch.zeros_like(kx)
    kmat = torch.stack((zero, -kz, ky, kz, zero, -kx, -ky, kx, zero), dim=1).reshape(b, 3, 3)
    eye = torch.eye(3, dtype=axis.dtype).expand(b, 3, 3)
    s = torch.sin(angle).reshape(b, 1, 1)
    c = torch.cos(angle).reshape(b, 1, 1)
    return eye + s * kmat + (1.0 - c) * (kmat @ kmat)


@dataclass
class Batch:
    pcl: torch.Tensor  # [B, N_o, 3] fp32 observed cloud (camera frame)
    prior: torch.Tensor  # [B, N_p, 3] fp32 normalised category prior
    init_pose: torch.Tensor  # [B, 3, 4] fp32
    init_scale: torch.Tensor  # [B, 3] fp32
    K: torch.Tensor  # [B, 3, 3] fp32
    obj_cls: torch.Tensor  # [B] int64

    def to(self, device) -> "Batch":
        return Batch(*(getattr(self, f).to(device) for f in ("pcl", "prior", "init_pose", "init_scale", "K", "obj_cls")))


def make_batch(batch: int, n_pts: int, seed: int, round_robin_cls: bool = False,
               fixtures: Optional[Fixtures] = None, n_prior: Optional[int] = None) -> Batch:
    """SURVEY.md 8(d) "Synthetic inputs".  Everything is drawn from one seeded CPU generator in
    float64 and cast to fp32 at the end.  ``n_prior``: prior points per object when it differs from the observed count."""
    b = _draw(batch, n_pts, seed, round_robin_cls, fixtures)[0]
    if n_prior is not None and n_prior != n_pts:
        fx = fixtures or load_fixtures()
        b.prior = resample_prior(fx.priors[b.obj_cls], n_prior).float().contiguous()
    return b


@dataclass
class TrainTargets:
    """Ground truth of a synthetic batch (the pose the observed cloud was rendered from), as the training
    forward takes it (reference: core/catre/engine/engine.py:305-318)."""
    gt_pose: torch.Tensor  # [B, 3, 4] fp32 = batch["obj_pose"]
    gt_scale: torch.Tensor  # [B, 3] fp32 = batch["obj_scale"]
    sym_y: torch.Tensor  # [B] bool: category symmetric about y (reference: ref/nocs.py:138-158, mug handle visible)


SYM_Y_CATEGORIES = ("bottle", "bowl", "can")


def make_train_batch(batch: int, n_pts: int, seed: int, round_robin_cls: bool = False,
                     fixtures: Optional[Fixtures] = None):
    """make_batch plus the ground truth it was rendered from: (Batch, TrainTargets).  Same random stream as
    make_batch, so the inputs are identical for a given seed."""
    b, rot_true, t_true, s_true = _draw(batch, n_pts, seed, round_robin_cls, fixtures)
    sym = torch.tensor([CATEGORIES[int(c)] in SYM_Y_CATEGORIES for c in b.obj_cls], dtype=torch.bool)
    gt_pose = torch.cat((rot_true, t_true.unsqueeze(2)), dim=2).float().contiguous()
    return b, TrainTargets(gt_pose=gt_pose, gt_scale=s_true.float().contiguous(), sym_y=sym)


def _draw(batch: int, n_pts: int, seed: int, round_robin_cls: bool, fixtures: Optional[Fixtures]):
    fx = fixtures or load_fixtures()
    g = torch.Generator().manual_seed(seed)
    n_inst = fx.init_pose.shape[0]
    if round_robin_cls:
        # config 5: force category b % 6, drawing without replacement inside each category
        idx = torch.empty(batch, dtype=torch.int64)
        for c in range(min(len(CATEGORIES), batch)):
            slots = torch.arange(c, batch, len(CATEGORIES))
            pool = torch.nonzero(fx.obj_cls == c).flatten()
            pick = pool[torch.randperm(pool.numel(), generator=g)[: slots.numel()]]
            idx[slots] = pick
    else:
        reps = (batch + n_inst - 1) // n_inst
        idx = torch.cat([torch.randperm(n_inst, generator=g) for _ in range(reps)])[:batch]
    pose0 = fx.init_pose[idx]
    scale0 = fx.init_scale[idx]
    cls = fx.obj_cls[idx]
    prior_full = fx.priors[cls]  # [B, 1024, 3] float64

    axis = torch.randn(batch, 3, generator=g, dtype=torch.float64)
    axis = axis / axis.norm(dim=1, keepdim=True).clamp_min(1e-12)
    angle = torch.randn(batch, generator=g, dtype=torch.float64) * math.radians(5.0)
    rot_true = _axis_angle_to_mat(axis, angle) @ pose0[:, :, :3]
    t_true = pose0[:, :, 3] + 0.01 * torch.randn(batch, 3, generator=g, dtype=torch.float64)
    s_true = scale0 * (1.0 + 0.05 * torch.randn(batch, 3, generator=g, dtype=torch.float64))

    cloud = (prior_full * s_true.unsqueeze(1)) @ rot_true.transpose(1, 2) + t_true.unsqueeze(1)  # [B,1024,3]
    view = -t_true / t_true.norm(dim=1, keepdim=True).clamp_min(1e-12)
    facing = ((cloud - t_true.unsqueeze(1)) * view.unsqueeze(1)).sum(-1) > -0.02  # [B,1024]
    pcl = torch.empty(batch, n_pts, 3, dtype=torch.float64)
    for b in range(batch):
        keep = torch.nonzero(facing[b]).flatten()
        if keep.numel() == 0:
            keep = torch.arange(cloud.shape[1])
        sel = keep[torch.randint(0, keep.numel(), (n_pts,), generator=g)]
        pcl[b] = cloud[b, sel]
    pcl = pcl + 0.002 * torch.randn(batch, n_pts, 3, generator=g, dtype=torch.float64)

    out = Batch(
        pcl=pcl.float().contiguous(),
        prior=resample_prior(prior_full, n_pts).float().contiguous(),
        init_pose=pose0.float().contiguous(),
        init_scale=scale0.float().contiguous(),
        K=torch.from_numpy(NOCS_REAL_K).expand(batch, 3, 3).contiguous(),
        obj_cls=cls.contiguous(),
    )
    return out, rot_true, t_true, s_true


def known_answer_inputs(fixtures: Optional[Fixtures] = None) -> Batch:
    """The known-answer case of SURVEY.md 8(c): mug prior, R=I, t=(0,0,1), s=(0.146,0.083,0.114),
    obs = Ry(10 deg) (1.05 s * P) + (0.01,-0.01,1.02), computed in float64 then cast."""
    fx = fixtures or load_fixtures()
    prior = fx.priors[CATEGORIES.index("mug")]
    s = torch.tensor([0.146, 0.083, 0.114], dtype=torch.float64)
    a = math.radians(10.0)
    ry = torch.tensor([[math.cos(a), 0.0, math.sin(a)], [0.0, 1.0, 0.0], [-math.sin(a), 0.0, math.cos(a)]],
                      dtype=torch.float64)
    obs = (1.05 * s * prior) @ ry.T + torch.tensor([0.01, -0.01, 1.02], dtype=torch.float64)
    pose = torch.cat((torch.eye(3, dtype=torch.float64), torch.tensor([[0.0], [0.0], [1.0]], dtype=torch.float64)), 1)
    return Batch(
        pcl=obs.float().unsqueeze(0).contiguous(),
        prior=prior.float().unsqueeze(0).contiguous(),
        init_pose=pose.float().unsqueeze(0).contiguous(),
        init_scale=s.float().unsqueeze(0).contiguous(),
        K=torch.from_numpy(NOCS_REAL_K).unsqueeze(0).contiguous(),
        obj_cls=torch.tensor([CATEGORIES.index("mug")], dtype=torch.int64),
    )
