"""Observed-cloud producer on the GPU: ``instances.pcl`` for all objects of one image.

Stands in for the per-object CPU loop of the reference's test data loader
(core/catre/datasets/data_loader.py:773-799 under the shipped config SAMPLE_DEPTH_FROM_BALL=True,
DEPTH_SAMPLE_BALL_RATIO=0.6, FPS_SAMPLE=False, OCCLUDE_MASK_TEST=False):
``depth_bp = backproject_th(depth, K)`` (lib/pysixd/misc.py:360-378) followed per object by
``crop_ball_from_depth_image(image, depth_bp, mask, pose, scale, ratio, K, num_points=NUM_PCL)``
(core/utils/cat_data_utils.py:380-400).  The per-pixel work (back-projection, mask & depth test, distances, radius
growth, ordered compaction, final gather) runs in libcatre_b200.so (csrc/cloud_kernels.cuh) for all objects at
once; the random draw is ``torch.randperm`` on the global CPU generator, called once per object in instance order
exactly as the reference does, so ``torch.manual_seed(s)`` reproduces the reference's cloud bit for bit.

No CPU fallback: needs the CUDA library and a device.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import engine as _engine

_SCRATCH: dict = {}
N_RADII = 10  # crop_ball_from_pts tries the initial radius and up to 9 enlargements (cat_data_utils.py:286-291)


def ball_radii(poses: torch.Tensor, scales: torch.Tensor, ratio: float) -> torch.Tensor:
    """[B, 10] fp32: the radii the reference would try for each object.  r0 = max(ratio * |R s|, 0.05)
    (cat_data_utils.py:386, :285), then r *= 1.10 per retry.  The reference keeps r as an fp32 tensor when the
    object term wins (fp32 multiplies) and as a Python float when the 0.05 floor wins (double multiplies, rounded
    to fp32 only in the comparison); both roundings are reproduced."""
    P, S = poses.detach().cpu().float(), scales.detach().cpu().float()
    # one mv + norm per object (the reference's own ops, so the fp32 rounding is the same), the x1.10 chain vectorised
    r0 = torch.stack([ratio * torch.norm(P[b, :, :3] @ S[b]) for b in range(P.shape[0])]) if P.shape[0] else torch.empty(0)
    cols = [r0]
    for _ in range(N_RADII - 1):
        cols.append(cols[-1] * 1.10)  # fp32 tensor times Python float: an fp32 multiply, like `radius *= 1.10`
    out = torch.stack(cols, dim=1)
    floor = r0 < 0.05  # Python's max(radius, 0.05) takes the float exactly when `0.05 > radius`
    if bool(floor.any()):
        chain, rf = [], 0.05
        for _ in range(N_RADII):
            chain.append(rf)  # double arithmetic, rounded to fp32 only below (as in `distance <= radius`)
            rf *= 1.10
        out[floor] = torch.tensor(chain, dtype=torch.float32)
    return out


def _intr(K) -> "ctypes.Array":
    Kt = torch.as_tensor(K).detach().cpu().float()
    return (ctypes.c_float * 4)(float(Kt[0, 0]), float(Kt[1, 1]), float(Kt[0, 2]), float(Kt[1, 2]))


def select_ball_points(depth: torch.Tensor, K, masks: torch.Tensor, poses: torch.Tensor, scales: torch.Tensor,
                       ratio: float = 0.6) -> Tuple[torch.Tensor, torch.Tensor]:
    """Device part 1: per object the ordered pixel ids inside the chosen ball.  depth [H,W] fp32 and masks [B,H,W]
    bool/uint8 may live on the CPU or the GPU; poses [B,3,4], scales [B,3].  Returns (sel_pix [B, H*W] int32 CUDA,
    n_sel [B] int32 CUDA); only the first n_sel[b] entries of row b are meaningful."""
    lib = _engine.load_library()
    if not torch.cuda.is_available():
        raise _engine.CatreError("catre_b200.cloud runs on CUDA only; there is no CPU path")
    dev = depth.device if depth.is_cuda else torch.device("cuda", torch.cuda.current_device())
    depth_d = depth.to(dev, torch.float32).contiguous()
    H, W = depth_d.shape
    masks_d = masks.to(dev).to(torch.uint8).contiguous()
    B = masks_d.shape[0]
    if tuple(masks_d.shape) != (B, H, W) or tuple(poses.shape) != (B, 3, 4) or tuple(scales.shape) != (B, 3):
        raise _engine.CatreError("shape mismatch between depth, masks, poses and scales")
    centers = poses[:, :, 3].detach().float().contiguous().to(dev)
    radii = ball_radii(poses, scales, ratio).to(dev)
    sel_pix = torch.empty((B, H * W), dtype=torch.int32, device=dev)
    n_sel = torch.zeros((B,), dtype=torch.int32, device=dev)
    key = (B, H, W, str(dev))
    scratch = _SCRATCH.get(key)
    if scratch is None:  # reused across calls (stream-ordered, so back-to-back calls on one stream are safe)
        scratch = _SCRATCH[key] = torch.empty((lib.catre_cloud_scratch_bytes(B, H, W),), dtype=torch.uint8, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    rc = lib.catre_cloud_select(depth_d.data_ptr(), masks_d.data_ptr(), _intr(K), centers.data_ptr(), radii.data_ptr(), N_RADII,
                                B, H, W, sel_pix.data_ptr(), n_sel.data_ptr(), scratch.data_ptr(), stream)
    if rc != 0:
        raise _engine.CatreError(f"catre_cloud_select failed ({rc}): {lib.catre_last_error(None).decode()}")
    return sel_pix, n_sel


def sample_object_clouds_batch(items: Sequence[Tuple[torch.Tensor, object, torch.Tensor, torch.Tensor, torch.Tensor]],
                               num_points: int = 1024, ratio: float = 0.6,
                               generator: Optional[torch.Generator] = None) -> List[torch.Tensor]:
    """``test_insts.pcl`` for SEVERAL images per call: ``items`` = [(depth, K, masks, poses, scales), ...] in loader order;
    returns one [B_i, num_points, 3] fp32 CUDA tensor per image.

    One call was host-bound (the kernels take a few microseconds; the device->host read of the point counts, the sample
    upload and a dozen small torch calls per image dominate), so the per-image costs that do not depend on the image are paid
    once per call here: every image's selection kernels are enqueued first, then ONE device->host copy brings all counts, the
    random draws run on the host in the reference's order -- image by image, object by object, one ``torch.randperm`` each on
    the same generator, so ``torch.manual_seed(s)`` still reproduces the reference's clouds bit for bit -- and ONE pinned
    upload carries every object's sample indices."""
    lib = _engine.load_library()
    sel = [select_ball_points(depth, K, masks, poses, scales, ratio) for depth, K, masks, poses, scales in items]
    if not sel:
        return []
    dev = sel[0][0].device
    counts = torch.cat([n for _, n in sel]).cpu().tolist()  # the one host sync: the draws below depend on the counts
    sample = torch.empty((len(counts), num_points), dtype=torch.int64, pin_memory=True)
    for b, n in enumerate(counts):  # loader order, instance order, one randperm per object: the reference's RNG call sequence
        if n == 0:
            raise ValueError(f"object {b} of the call has no valid depth pixel under its mask")
        length = n
        while length < num_points:  # `while len(idx) < num_points: idx = cat([idx, idx])` (cat_data_utils.py:297-298)
            length *= 2
        sample[b] = torch.randperm(length, generator=generator)[:num_points]
    sample_d = sample.to(dev, non_blocking=True)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    out, lo = [], 0
    for (depth, K, masks, poses, scales), (sel_pix, n_sel) in zip(items, sel):
        B = int(n_sel.shape[0])
        H, W = depth.shape
        depth_d = depth.to(dev, torch.float32).contiguous()
        pcl = torch.empty((B, num_points, 3), dtype=torch.float32, device=dev)
        rc = lib.catre_cloud_gather(depth_d.data_ptr(), _intr(K), sel_pix.data_ptr(), n_sel.data_ptr(),
                                    sample_d[lo:lo + B].data_ptr(), B, H, W, num_points, pcl.data_ptr(), stream)
        if rc != 0:
            raise _engine.CatreError(f"catre_cloud_gather failed ({rc}): {lib.catre_last_error(None).decode()}")
        out.append(pcl)
        lo += B
    return out


def sample_object_clouds(depth: torch.Tensor, K, masks: torch.Tensor, poses: torch.Tensor, scales: torch.Tensor,
                         num_points: int = 1024, ratio: float = 0.6,
                         generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """``test_insts.pcl`` for one image: [B, num_points, 3] fp32 on the GPU (data_loader.py:773-799)."""
    return sample_object_clouds_batch([(depth, K, masks, poses, scales)], num_points, ratio, generator)[0]
