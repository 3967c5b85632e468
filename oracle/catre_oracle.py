"""CPU oracle for the CATRE iterative pose-refinement hot path.

TEST INFRASTRUCTURE ONLY.  This is a CPU restatement (torch, CPU tensors, fp32 or fp64) of the
reference's algorithm for the path in SURVEY.md section 8.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and only as the checker / the CPU baseline, never as the product path.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified reference from
``/root/reference`` (through an import shim for its missing third-party packages), runs it on the
shipped checkpoint and writes the golden vectors under ``tests/golden/``;
``tests/test_oracle.py`` checks this restatement against those vectors and against the
known-answer vector recorded in SURVEY.md section 8(c).

Each function cites the reference file:line it follows.  Nothing here is copied from the
reference: it is written against the semantic spec in SURVEY.md appendix A, with plain functional
torch ops on a flat ``{checkpoint-name: tensor}`` weight dict.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]

# REAL275 camera intrinsics (reference: ref/nocs.py:103)
NOCS_REAL_K = ((591.0125, 0.0, 322.525), (0.0, 590.16775, 244.11084), (0.0, 0.0, 1.0))

GN_GROUPS = 32  # configs/catre/NOCS_REAL/aug05_..._120e.py:92,109 (num_gn_groups)
GN_EPS = 1e-5  # torch.nn.GroupNorm default (lib/torch_utils/layers/layer_utils.py:51)


def _pw(w: Weights, name: str, x: torch.Tensor) -> torch.Tensor:
    """Point-wise (kernel 1) Conv1d on [B, C, N]."""
    return F.conv1d(x, w[name + ".weight"], w[name + ".bias"])


def _fc(w: Weights, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, w[name + ".weight"], w[name + ".bias"])


def tnet(w: Weights, prefix: str, x: torch.Tensor, k: int) -> torch.Tensor:
    """STN3d (k=3) / STNkd (k=64): core/catre/models/pointnets/pointnet.py:24-41, 57-78.

    x: [S, k, N] -> [S, k, k] = I + fc3(relu(fc2(relu(fc1(max_n relu(c3(relu(c2(relu(c1 x))))))))))
    """
    h = F.relu(_pw(w, prefix + ".conv1", x))
    h = F.relu(_pw(w, prefix + ".conv2", h))
    h = F.relu(_pw(w, prefix + ".conv3", h))
    h = torch.max(h, 2)[0]
    h = F.relu(_fc(w, prefix + ".fc1", h))
    h = F.relu(_fc(w, prefix + ".fc2", h))
    h = _fc(w, prefix + ".fc3", h)
    eye = torch.eye(k, dtype=h.dtype, device=h.device).reshape(1, k * k)
    return (h + eye).reshape(-1, k, k)


def pointnet_feat(w: Weights, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """PointNetfeat.forward with global_feat=False, feature_transform=True
    (core/catre/models/pointnets/pointnet.py:97-121).

    x: [S, 3, N].  Returns (g [S, 1024], pointfeat [S, 64, N]); the reference's output is
    ``cat([g.repeat(N), pointfeat], dim=1)`` which callers here never need materialised.
    """
    t3 = tnet(w, "pcl_net.stn", x, 3)
    x = torch.bmm(x.transpose(2, 1), t3).transpose(2, 1)  # x'_j = sum_i x_i T3[i, j]
    h1 = F.relu(_pw(w, "pcl_net.conv1", x))
    t64 = tnet(w, "pcl_net.fstn", h1, 64)
    pf = torch.bmm(h1.transpose(2, 1), t64).transpose(2, 1)
    h = F.relu(_pw(w, "pcl_net.conv2", pf))
    h = F.relu(_pw(w, "pcl_net.conv3", h))
    h = _pw(w, "pcl_net.conv4", h)  # no ReLU after conv4 (pointnet.py:114)
    g = torch.max(h, 2)[0]
    return g, pf


def ts_head(w: Weights, feat: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """FC_TransSizeHead.forward: core/catre/models/heads/fc_trans_size_head.py:61-70."""
    h = _fc(w, "ts_head.linears.0", feat)
    h = F.gelu(F.group_norm(h, GN_GROUPS, w["ts_head.linears.1.weight"], w["ts_head.linears.1.bias"], GN_EPS))
    h = _fc(w, "ts_head.linears.3", h)
    h = F.gelu(F.group_norm(h, GN_GROUPS, w["ts_head.linears.4.weight"], w["ts_head.linears.4.bias"], GN_EPS))
    return _fc(w, "ts_head.fc_t", h), _fc(w, "ts_head.fc_s", h)


def rot_head_one(w: Weights, prefix: str, feat: torch.Tensor) -> torch.Tensor:
    """RotHead.forward: core/catre/models/heads/conv_out_per_rot_head.py:126-140.

    feat: [B, 1088, P] -> [B, 3] (learned weighted sum over the point index via conv_p).
    """
    h = _pw(w, prefix + ".layers.0", feat)
    h = F.gelu(F.group_norm(h, GN_GROUPS, w[prefix + ".layers.1.weight"], w[prefix + ".layers.1.bias"], GN_EPS))
    h = _pw(w, prefix + ".layers.3", h)
    h = F.gelu(F.group_norm(h, GN_GROUPS, w[prefix + ".layers.4.weight"], w[prefix + ".layers.4.bias"], GN_EPS))
    h = _pw(w, prefix + ".neck.0", h)  # [B, 3, P]
    h = F.conv1d(h.permute(0, 2, 1), w[prefix + ".conv_p.weight"], w[prefix + ".conv_p.bias"])  # [B, 1, 3]
    return h.squeeze(1)


def rot6d_to_mat(d6: torch.Tensor) -> torch.Tensor:
    """Gram-Schmidt of the 6-D representation: core/utils/rot_reps.py:34-55 (columns x, y, z)."""
    x = F.normalize(d6[..., 0:3], p=2, dim=-1)
    z = F.normalize(torch.cross(x, d6[..., 3:6], dim=-1), p=2, dim=-1)
    y = torch.cross(z, x, dim=-1)
    return torch.stack((x, y, z), dim=-1)


def pose_update(
    d_rot: torch.Tensor, d_t: torch.Tensor, d_s: torch.Tensor,
    rot: torch.Tensor, t: torch.Tensor, s: torch.Tensor, K: torch.Tensor,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """pose_scale_from_delta_init with the shipped settings (image space, K aware, cosypose z,
    additive scale, ego rotation): core/catre/models/pose_scale_from_delta_init.py:47-95."""
    z_src = t[:, 2:3]
    z_tgt = d_t[:, 2:3] * z_src
    fxfy = torch.stack((K[:, 0, 0], K[:, 1, 1]), dim=1)
    xy_tgt = z_tgt * (d_t[:, :2] / fxfy + t[:, :2] / z_src)
    return d_rot @ rot, torch.cat((xy_tgt, z_tgt), dim=-1), s + d_s


def forward_once(
    w: Weights, x: torch.Tensor, tfd_kps: torch.Tensor,
    init_pose: torch.Tensor, init_scale: torch.Tensor, K: torch.Tensor,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """One refinement iteration = CATRE_disR_shared.forward(do_loss=False)
    (core/catre/models/CATRE_disR_shared.py:57-124) under the shipped config
    (WITH_KPS_FEATURE=False, WITH_INIT_SCALE=True, CLASS_AWARE=False, REFINE_SCLAE=True).

    x [B,3,N_o], tfd_kps [B,3,N_p], init_pose [B,3,4], init_scale [B,3], K [B,3,3]
    -> pose [B,3,4], scale [B,3]
    """
    g_o, pf_o = pointnet_feat(w, x)
    g_p, pf_p = pointnet_feat(w, tfd_kps)
    n_o, n_p = x.shape[2], tfd_kps.shape[2]
    f_o = torch.cat((g_o, torch.max(pf_o, 2)[0]), dim=1)  # max_n of cat([g.repeat, pf]) (:69)
    d_t, d_s = ts_head(w, torch.cat((f_o, init_scale), dim=1))
    feat_o = torch.cat((g_o.unsqueeze(2).expand(-1, -1, n_o), pf_o), dim=1)
    feat_p = torch.cat((g_p.unsqueeze(2).expand(-1, -1, n_p), pf_p), dim=1)
    rot_feat = torch.cat((feat_o, feat_p), dim=2)  # points: obs first, prior second (:86)
    r6 = torch.cat(
        (rot_head_one(w, "rot_head.rot_head_x", rot_feat), rot_head_one(w, "rot_head.rot_head_y", rot_feat)), dim=1
    )
    rot, t, s = pose_update(rot6d_to_mat(r6), d_t, d_s, init_pose[:, :3, :3], init_pose[:, :3, 3], init_scale, K)
    return torch.cat((rot, t.reshape(-1, 3, 1)), dim=-1), s


def update_points(
    pcl: torch.Tensor, kps: torch.Tensor, pose: torch.Tensor, scale: torch.Tensor
) -> Tuple[torch.Tensor, torch.Tensor]:
    """batch_updater_test with ZERO_CENTER_INPUT=True, KPS_TYPE="mean_shape"
    (core/catre/engine/batch_test.py:78-97, lib/pysixd/misc.py:1011-1026).

    pcl [B,N_o,3], kps [B,N_p,3] -> x [B,3,N_o] = pcl - t, tfd_kps [B,3,N_p] = R (s * kps)
    """
    b, n_p = kps.shape[0], kps.shape[1]
    rot, t = pose[:, :3, :3], pose[:, :3, 3]
    scaled = kps * scale.unsqueeze(1)
    tfd = (rot.reshape(b, 1, 3, 3) @ scaled.reshape(b, n_p, 3, 1)).squeeze(-1)
    return pcl.permute(0, 2, 1) - t.reshape(b, 3, 1), tfd.permute(0, 2, 1)


@torch.no_grad()
def refine(
    w: Weights, pcl: torch.Tensor, kps: torch.Tensor,
    init_pose: torch.Tensor, init_scale: torch.Tensor, K: torch.Tensor, n_iter: int,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """The evaluator's K-loop (core/catre/engine/catre_evaluator.py:292-311).

    Returns every iteration's pose, like the evaluator's out_dict:
    poses [n_iter+1, B, 3, 4] and scales [n_iter+1, B, 3]; entry 0 is the initial pose.
    """
    poses, scales = [init_pose], [init_scale]
    for _ in range(n_iter):
        x, tfd = update_points(pcl, kps, poses[-1], scales[-1])
        p, s = forward_once(w, x, tfd, poses[-1], scales[-1], K)
        poses.append(p)
        scales.append(s)
    return torch.stack(poses), torch.stack(scales)


def cast_weights(w: Weights, dtype: torch.dtype) -> Weights:
    return {k: v.to(dtype) for k, v in w.items()}


def resize_conv_p(w: Weights, n_pts: int, n_prior=None) -> Weights:
    """Fixture-defined weights for N != 1024 (SURVEY.md 8(d) "Weights per config"): everything
    from the checkpoint except the two conv_p.weight [1, 2*1024, 1], whose obs half and prior half
    are each linearly re-sampled to n_pts and scaled by 1024 / n_pts."""
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
