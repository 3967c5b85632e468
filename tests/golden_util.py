"""Helpers shared by the parity tests: load a committed golden case (inputs + reference outputs)."""
import json
import os
from dataclasses import dataclass

import numpy as np
import torch

from catre_b200 import synth

GOLDEN_DIR = synth.GOLDEN_DIR

# Parity tolerance from BASELINE.json north_star: every (R, t, s) component within 1e-4 (fp32).
TOL = 1e-4


@dataclass
class GoldenCase:
    name: str
    batch: synth.Batch
    poses: torch.Tensor  # [K+1, B, 3, 4] reference output
    scales: torch.Tensor  # [K+1, B, 3]
    n_iter: int
    n_pts: int


def index():
    with open(os.path.join(GOLDEN_DIR, "golden_index.json")) as f:
        return json.load(f)


def case_names():
    return list(index()["cases"].keys())


def load_case(name: str) -> GoldenCase:
    meta = index()["cases"][name]
    z = np.load(os.path.join(GOLDEN_DIR, f"golden_{name}.npz"))
    fx = synth.load_fixtures()
    cls = torch.from_numpy(z["prior_cls"].astype(np.int64))
    prior = synth.resample_prior(fx.priors[cls], meta["n_pts"]).float().contiguous()
    batch = synth.Batch(
        pcl=torch.from_numpy(z["pcl"]), prior=prior, init_pose=torch.from_numpy(z["init_pose"]),
        init_scale=torch.from_numpy(z["init_scale"]), K=torch.from_numpy(z["K"]), obj_cls=cls,
    )
    return GoldenCase(name, batch, torch.from_numpy(z["poses"]), torch.from_numpy(z["scales"]),
                      meta["n_iter"], meta["n_pts"])


def max_abs_err(poses, scales, ref_poses, ref_scales):
    """max |delta| over R, t and s components (all iterations)."""
    poses, scales = poses.double().cpu(), scales.double().cpu()
    ref_poses, ref_scales = ref_poses.double(), ref_scales.double()
    e_r = (poses[..., :3] - ref_poses[..., :3]).abs().max().item()
    e_t = (poses[..., 3] - ref_poses[..., 3]).abs().max().item()
    e_s = (scales - ref_scales).abs().max().item()
    return e_r, e_t, e_s
