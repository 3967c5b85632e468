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


# ---- full-size cases (tests/golden/make_golden_full.py): outputs committed, inputs regenerated from the seed ----
# Per-mode REGRESSION thresholds beside the 1e-4 contract ("tighten the gate"): distance to the fp64
# oracle on the same inputs, i.e. free of the reference's own fp32 noise.  Keyed by (precision, n_iter).
# (The reference's OWN fp32 run is 4.1e-6 / 4.6e-6 from the fp64 oracle on the 384-object K=4 / 256-object K=8 cases --
# golden_full_index.json "ref_fp32_vs_fp64_max_abs" -- so the fp32 mode's gate sits at ~2.5x that noise, not below it.)
# f16x3 at K = 8: measured 5.0e-5 on the 256-object N=2048 case (profiles/r02a_pytest_gpu.log), i.e. AT the 5e-5 the
# round-1 verdict proposed; the gate is 7.5e-5 -- a 1.5x regression still fails, the 1e-4 contract keeps 25 % of margin.
REGRESSION_TOL = {("fp32", 4): 1e-5, ("f16x3", 4): 2e-5, ("fp32", 8): 2e-5, ("f16x3", 8): 7.5e-5}


@dataclass
class FullCase:
    name: str
    batch: synth.Batch
    poses: torch.Tensor    # reference fp32 output (the pin)
    scales: torch.Tensor
    poses64: torch.Tensor  # fp64 oracle on the same inputs (the yardstick)
    scales64: torch.Tensor
    n_iter: int
    n_pts: int
    meta: dict


def full_index():
    with open(os.path.join(GOLDEN_DIR, "golden_full_index.json")) as f:
        return json.load(f)


def full_case_names():
    try:
        return list(full_index()["cases"].keys())
    except FileNotFoundError:
        return []


def batch_digest(batch: synth.Batch) -> str:
    import hashlib

    h = hashlib.sha256()
    for f in ("pcl", "prior", "init_pose", "init_scale", "K"):
        h.update(getattr(batch, f).contiguous().numpy().tobytes())
    h.update(batch.obj_cls.numpy().astype(np.int64).tobytes())
    return h.hexdigest()


def load_full_case(name: str, check_digest: bool = True) -> FullCase:
    meta = full_index()["cases"][name]
    z = np.load(os.path.join(GOLDEN_DIR, f"golden_full_{name}.npz"))
    batch = synth.make_batch(meta["batch"], meta["n_pts"], meta["seed"], meta["round_robin"])
    if check_digest:
        assert batch_digest(batch) == meta["input_sha256"], f"{name}: the seeded generator no longer reproduces the golden inputs"
    return FullCase(name, batch, torch.from_numpy(z["poses"]), torch.from_numpy(z["scales"]), torch.from_numpy(z["poses64"]),
                    torch.from_numpy(z["scales64"]), meta["n_iter"], meta["n_pts"], meta)


def max_abs_err_nan_aware(poses, scales, ref_poses, ref_scales):
    """As max_abs_err, but positions where the reference is NaN must be NaN in the result too (and nowhere else);
    they are excluded from the maximum.  (One REAL275 initial pose has t = 0; the reference returns NaN for it.)"""
    poses, scales = poses.double().cpu(), scales.double().cpu()
    ref_poses, ref_scales = ref_poses.double(), ref_scales.double()
    assert torch.equal(torch.isnan(poses), torch.isnan(ref_poses)), "NaN pattern of the poses differs from the reference's"
    assert torch.equal(torch.isnan(scales), torch.isnan(ref_scales)), "NaN pattern of the scales differs from the reference's"
    dp = torch.nan_to_num((poses - ref_poses).abs(), nan=0.0)
    ds = torch.nan_to_num((scales - ref_scales).abs(), nan=0.0)
    return dp[..., :3].max().item(), dp[..., 3].max().item(), ds.max().item()
