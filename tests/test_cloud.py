"""Observed-cloud producer: the CPU oracle against golden vectors made by the unmodified
reference functions (tests/golden/make_golden_cloud.py), and -- on the GPU -- the CUDA producer against both.
Bit-exact throughout: this is index / IEEE-fp32 work, and the random draw uses the same torch.randperm calls."""
import os

import numpy as np
import pytest
import torch

from catre_b200 import cloud
from oracle import cloud_oracle as co

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_cloud.npz")


def load_scene(si):
    z = np.load(GOLDEN)
    depth = torch.from_numpy(z[f"depth_{si}"])
    W = depth.shape[1]
    masks = torch.from_numpy(np.unpackbits(z[f"masks_{si}"], axis=-1)[..., :W].astype(bool))
    return (depth, z["K"], masks, torch.from_numpy(z[f"poses_{si}"]), torch.from_numpy(z[f"scales_{si}"]), int(z[f"seed_{si}"]),
            torch.from_numpy(z[f"pcl_{si}"]), z)


@pytest.mark.parametrize("si", [0, 1])
def test_oracle_matches_reference_golden(si):
    depth, K, masks, poses, scales, seed, ref_pcl, z = load_scene(si)
    torch.manual_seed(seed)
    pcl = co.sample_clouds(depth, K, masks, poses, scales, 1024)
    assert torch.equal(pcl, ref_pcl)
    if si == 0:
        assert torch.equal(co.backproject(depth, K), torch.from_numpy(z["depth_bp_0"]))


def test_oracle_branches_are_exercised():
    """The scene was built to hit every branch of crop_ball_from_pts: radius floor, growth, empty ball, doubling."""
    depth, K, masks, poses, scales, _, _, _ = load_scene(0)
    bp = co.backproject(depth, K)
    sizes, floors = [], []
    for m, p, s in zip(masks, poses, scales):
        radii = co.ball_radii(p, s, 0.6)
        floors.append(not isinstance(radii[0], torch.Tensor))
        valid = torch.logical_and(m, bp[:, :, 2] > 0).flatten().nonzero().squeeze(1)
        pts = bp.reshape(-1, 3)[valid]
        d = torch.sqrt(((pts - p[:, 3]) ** 2).sum(-1))
        first = int((d <= radii[0]).sum())
        sel = co.ball_indices(pts, p[:, 3], radii)
        sizes.append((len(valid), first, len(sel)))
    assert floors == [False, True, False, False, False]
    assert sizes[0][2] >= 1024                       # no doubling needed
    assert sizes[2][1] < 10 <= sizes[2][2]           # radius had to grow
    assert sizes[3][1] == 0 and sizes[3][2] == sizes[3][0]  # empty ball after 10 tries -> all valid points
    assert sizes[4][2] < 1024                        # index doubling
    radii = cloud.ball_radii(poses, scales, 0.6)     # the product's host-side radius rule == the oracle's
    for b in range(poses.shape[0]):
        want = torch.tensor([float(torch.tensor(r, dtype=torch.float32)) if not isinstance(r, torch.Tensor) else float(r)
                             for r in co.ball_radii(poses[b], scales[b], 0.6)])
        assert torch.equal(radii[b], want)


def test_oracle_rejects_object_without_depth():
    depth, K, masks, poses, scales, _, _, _ = load_scene(0)
    empty = torch.zeros_like(masks[0])
    with pytest.raises(ValueError):
        co.sample_cloud(co.backproject(depth, K), empty, poses[0], scales[0], 1024)


@pytest.mark.gpu
@pytest.mark.parametrize("si", [0, 1])
def test_cuda_producer_matches_reference_golden(si):
    depth, K, masks, poses, scales, seed, ref_pcl, _ = load_scene(si)
    torch.manual_seed(seed)
    pcl = cloud.sample_object_clouds(depth, K, masks, poses, scales, 1024)
    assert pcl.is_cuda and torch.equal(pcl.cpu(), ref_pcl)
    # the selected index lists themselves (before the random draw) equal the oracle's
    sel_pix, n_sel = cloud.select_ball_points(depth, K, masks, poses, scales)
    bp = co.backproject(depth, K)
    for b in range(masks.shape[0]):
        valid = torch.logical_and(masks[b], bp[:, :, 2] > 0).flatten().nonzero().squeeze(1)
        sel = co.sample_cloud(bp, masks[b], poses[b], scales[b], 1024)[1]
        assert int(n_sel[b]) == len(sel)
        assert torch.equal(sel_pix[b, : len(sel)].cpu().long(), valid[sel])


@pytest.mark.gpu
def test_cuda_producer_full_resolution_vs_oracle():
    """REAL275 resolution (480 x 640), 6 objects, fresh seeded scene: CUDA producer == oracle, bit for bit."""
    g = torch.Generator().manual_seed(5)
    H, W = 480, 640
    K = np.array([[591.0125, 0, 322.525], [0, 590.16775, 244.11084], [0, 0, 1]], dtype=np.float32)
    v, u = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    depth = 1.2 + 0.0005 * v + 0.0003 * u + 0.002 * torch.randn(H, W, generator=g)
    depth[torch.rand(H, W, generator=g) < 0.1] = 0.0
    masks, poses, scales = [], [], []
    for i in range(6):
        cu, cv, r = 80 + 90 * i, 100 + 50 * i, 20 + 9 * i
        m = (u - cu) ** 2 + (v - cv) ** 2 <= r ** 2
        depth = torch.where(m & (depth > 0), depth - 0.4, depth)
        masks.append(m)
        z = 0.8 + 0.0005 * cv + 0.0003 * cu
        t = torch.tensor([(cu - K[0, 2]) * z / K[0, 0], (cv - K[1, 2]) * z / K[1, 1], z + (0.5 if i == 4 else 0.0)])
        poses.append(torch.cat((torch.eye(3), t.reshape(3, 1)), dim=1))
        scales.append(torch.tensor([0.004, 0.004, 0.004]) if i == 2 else torch.tensor([0.1, 0.15, 0.12]) * (1 + 0.2 * i))
    depth = depth.float().contiguous()
    masks, poses, scales = torch.stack(masks), torch.stack(poses).float(), torch.stack(scales).float()
    torch.manual_seed(77)
    want = co.sample_clouds(depth, K, masks, poses, scales, 1024)
    torch.manual_seed(77)
    got = cloud.sample_object_clouds(depth.cuda(), K, masks.cuda(), poses, scales, 1024)
    assert torch.equal(got.cpu(), want)
    with pytest.raises(ValueError):
        cloud.sample_object_clouds(depth, K, torch.zeros_like(masks[:1]), poses[:1], scales[:1], 1024)


@pytest.mark.gpu
def test_batched_call_equals_the_per_image_sequence():
    """Several images per call (one count read-back, one sample upload): the same clouds, bit for bit, as the per-image calls in
    loader order -- the host draws stay in the reference's order on the same generator."""
    scenes = [load_scene(0), load_scene(1), load_scene(0)]
    items = [(d, K, m, p, s) for d, K, m, p, s, *_ in scenes]
    torch.manual_seed(123)
    one_by_one = [cloud.sample_object_clouds(*it, 1024) for it in items]
    torch.manual_seed(123)
    batched = cloud.sample_object_clouds_batch(items, 1024)
    assert len(batched) == len(one_by_one)
    for a, b in zip(batched, one_by_one):
        assert torch.equal(a, b)
    # and against the CPU oracle run image by image with the same seed
    torch.manual_seed(123)
    for it, got in zip(items, batched):
        assert torch.equal(got.cpu(), co.sample_clouds(*it, 1024))
    assert cloud.sample_object_clouds_batch([], 1024) == []
