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
    out = gt_rot.detach().cpu().clone().numpy()
    pred = pred_rot.detach().cpu().numpy()
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
    return torch.from_numpy(out).to(dtype=gt_rot.dtype, device=gt_rot.device)


def catre_loss(rot: torch.Tensor, trans: torch.Tensor, scale: torch.Tensor, gt_rot: torch.Tensor, gt_trans: torch.Tensor,
               gt_scale: torch.Tensor, kps: torch.Tensor, sym_info: List[Optional[np.ndarray]],
               weights: Tuple[float, float, float, float] = (1.0, 1.0, 1.0, 1.0)) -> Dict[str, torch.Tensor]:
    """CATRE_disR_shared.catre_loss with the shipped LOSS_CFG (core/catre/models/CATRE_disR_shared.py:168-288;
    configs/catre/NOCS_REAL/aug05_..._120e.py:115-134: symmetric point-matching loss on R only with scale, L1;
    angular rotation loss for asymmetric objects, L1 on the y axis for symmetric ones; L1 on xy / z; L1 on scale;
    every weight 1).  ``weights`` = (PM_LW, ROT_LW, TRANS_LW, SCALE_LW), each multiplying its terms (:217, 238, 250, 262-263,
    286)."""
    w_pm, w_rot, w_trans, w_scale = weights
    loss: Dict[str, torch.Tensor] = {}
    # point matching (core/catre/losses/pm_loss.py:110-130): R (s * kps) vs R_gt* (s_gt * kps), L1 mean, times 3
    gt_sym = closest_sym_rot(rot, gt_rot, sym_info)
    est = (kps * scale.unsqueeze(1)) @ rot.transpose(1, 2)
    tgt = (kps * gt_scale.unsqueeze(1)) @ gt_sym.transpose(1, 2)
    loss["loss_PM_R"] = 3.0 * (est - tgt).abs().mean() * w_pm
    # rotation (CATRE_disR_shared.py:222-250; core/catre/losses/rot_loss.py:45-58)
    is_sym = torch.tensor([s is not None for s in sym_info], device=rot.device)
    if (~is_sym).any():
        m = rot[~is_sym] @ gt_rot[~is_sym].transpose(1, 2)
        cos = (m.diagonal(dim1=1, dim2=2).sum(1) - 1.0) / 2.0
        loss["loss_rot"] = ((1.0 - cos) / 2.0).mean() * w_rot
    if is_sym.any():
        loss["loss_yaxis_rot"] = (rot[is_sym][:, :, 1] - gt_rot[is_sym][:, :, 1]).abs().mean() * w_rot
    # translation, disentangled (CATRE_disR_shared.py:253-262) and scale (:277-286)
    loss["loss_trans_xy"] = (trans[:, :2] - gt_trans[:, :2]).abs().mean() * w_trans
    loss["loss_trans_z"] = (trans[:, 2] - gt_trans[:, 2]).abs().mean() * w_trans
    loss["loss_scale"] = (scale - gt_scale).abs().mean() * w_scale
    return loss


def train_step(w: Weights, pcl: torch.Tensor, kps: torch.Tensor, pose: torch.Tensor, scale: torch.Tensor, K: torch.Tensor,
               gt_pose: torch.Tensor, gt_scale: torch.Tensor, sym_info: List[Optional[np.ndarray]],
               weights: Tuple[float, float, float, float] = (1.0, 1.0, 1.0, 1.0)):
    """One refinement iteration of the training loop (core/catre/engine/engine.py:293-352 without the optimiser):
    re-pose the points with the current (detached) estimate, forward, losses, backward of their sum.
    Returns (pose', scale', loss dict of floats, {name: gradient})."""
    wg = {k: v.clone().requires_grad_(k not in UNUSED) for k, v in w.items()}
    x, tfd = co.update_points(pcl, kps, pose, scale)
    new_pose, new_scale = co.forward_once(wg, x, tfd, pose, scale, K)
    losses = catre_loss(new_pose[:, :3, :3], new_pose[:, :3, 3], new_scale, gt_pose[:, :3, :3], gt_pose[:, :3, 3], gt_scale,
                        kps, sym_info, weights)
    sum(losses.values()).backward()
    grads = {k: v.grad for k, v in wg.items() if v.grad is not None}
    return new_pose.detach(), new_scale.detach(), {k: float(v.detach()) for k, v in losses.items()}, grads


# ======================================================================================================
# Hand-derived backward in the CUDA chain's decomposition.  Layout as on the GPU: sets s = 0..2B-1 (set 2b = the
# observed cloud of object b, 2b+1 = its prior), per-point tensors point-major [S, N, C].
# ======================================================================================================
def _gelu_grad(x: torch.Tensor) -> torch.Tensor:
    """d/dx [0.5 x (1 + erf(x / sqrt 2))] = Phi(x) + x phi(x)."""
    return 0.5 * (1.0 + torch.erf(x * 0.7071067811865476)) + x * torch.exp(-0.5 * x * x) * 0.3989422804014327


def _gn_stats(y: torch.Tensor, red_dims) -> Tuple[torch.Tensor, torch.Tensor]:
    """Biased mean / rstd of a GroupNorm group (eps 1e-5); y already reshaped so that red_dims span one group."""
    mu = y.mean(dim=red_dims, keepdim=True)
    var = ((y - mu) ** 2).mean(dim=red_dims, keepdim=True)
    return mu, torch.rsqrt(var + co.GN_EPS)


def _gn_gelu_backward(dy_out, y, mu, rstd, gamma, beta, groups_view, red_dims):
    """Backward of u = gelu(gamma * (y - mu) * rstd + beta) given du (= dy_out), from two group sums:
    with xhat = (y - mu) rstd, g = du gelu'(.) gamma:  dy = rstd (g - mean_grp(g) - xhat mean_grp(g xhat)).
    `groups_view(t)` reshapes a tensor so that red_dims span one group.  Returns (dy, dgamma_elem, dbeta_elem) where
    the last two still have to be summed over everything but the channel."""
    xhat = (y - mu) * rstd
    dn = dy_out * _gelu_grad(gamma * xhat + beta)  # gradient at the normalised-and-affine value
    g = dn * gamma
    m1 = groups_view(g).mean(dim=red_dims, keepdim=True)
    m2 = groups_view(g * xhat).mean(dim=red_dims, keepdim=True)
    dy = rstd * (g - m1.reshape(mu.shape) - xhat * m2.reshape(mu.shape))
    return dy, dn * xhat, dn


def manual_forward(w: Weights, q: torch.Tensor, pose: torch.Tensor, scale: torch.Tensor, K: torch.Tensor) -> dict:
    """The forward of one iteration in the engine's layout, keeping what the backward needs.  q [S, N, 3] are the
    re-posed points (set 2b = pcl_b - t_b, set 2b+1 = R_b (s_b * kps_b))."""
    S, N, _ = q.shape
    B = S // 2
    sv: dict = {"q": q}

    def pw(name, x):  # per-point layer on [S, N, K]
        return x @ w[name + ".weight"][:, :, 0].t() + w[name + ".bias"]

    def fc(name, x):
        return x @ w[name + ".weight"].t() + w[name + ".bias"]

    def tnet(pre, x, k, tag):
        sv[tag + "64"] = c1 = F.relu(pw(pre + ".conv1", x))
        sv[tag + "128"] = c2 = F.relu(pw(pre + ".conv2", c1))
        z = F.relu(pw(pre + ".conv3", c2))  # [S, N, 1024]: never stored on the GPU
        sv[tag + "max"], sv[tag + "arg"] = z.max(dim=1)
        sv[tag + "fc1"] = f1 = F.relu(fc(pre + ".fc1", sv[tag + "max"]))
        sv[tag + "fc2"] = f2 = F.relu(fc(pre + ".fc2", f1))
        return (fc(pre + ".fc3", f2) + torch.eye(k, dtype=q.dtype).reshape(1, k * k)).reshape(S, k, k)

    sv["t3"] = t3 = tnet("pcl_net.stn", q, 3, "s")
    sv["qp"] = qp = q @ t3
    sv["h1"] = h1 = F.relu(pw("pcl_net.conv1", qp))
    sv["t64"] = t64 = tnet("pcl_net.fstn", h1, 64, "f")
    sv["pf"] = pf = h1 @ t64
    sv["a128"] = a128 = F.relu(pw("pcl_net.conv2", pf))
    sv["a512"] = a512 = F.relu(pw("pcl_net.conv3", a128))
    sv["g"], sv["garg"] = pw("pcl_net.conv4", a512).max(dim=1)
    sv["pfmax"], sv["pfarg"] = pf.max(dim=1)

    # ts head on the observed sets
    g_o, pfm_o = sv["g"][0::2], sv["pfmax"][0::2]
    sv["ts_in"] = f = torch.cat((g_o, pfm_o, scale), dim=1)
    sv["ts_y0"] = y0 = fc("ts_head.linears.0", f)
    sv["ts_st0"] = _gn_stats(y0.reshape(B, 32, 8), (2,))
    sv["ts_u0"] = u0 = F.gelu(((y0.reshape(B, 32, 8) - sv["ts_st0"][0]) * sv["ts_st0"][1]).reshape(B, 256)
                              * w["ts_head.linears.1.weight"] + w["ts_head.linears.1.bias"])
    sv["ts_y1"] = y1 = fc("ts_head.linears.3", u0)
    sv["ts_st1"] = _gn_stats(y1.reshape(B, 32, 8), (2,))
    sv["ts_u1"] = u1 = F.gelu(((y1.reshape(B, 32, 8) - sv["ts_st1"][0]) * sv["ts_st1"][1]).reshape(B, 256)
                              * w["ts_head.linears.4.weight"] + w["ts_head.linears.4.bias"])
    d_t, d_s = fc("ts_head.fc_t", u1), fc("ts_head.fc_s", u1)

    # rotation heads with the layer-0 split: W0 [g_set | pf_p] = W0[:, :1024] g_set + W0[:, 1024:] pf_p
    r6 = []
    pf_obj = pf.reshape(B, 2 * N, 64)  # an object's points: observed first, prior second
    for h, pre in enumerate(("rot_head.rot_head_x", "rot_head.rot_head_y")):
        w0 = w[pre + ".layers.0.weight"][:, :, 0]
        sv[f"cset{h}"] = cset = sv["g"] @ w0[:, :1024].t() + w[pre + ".layers.0.bias"]  # [S, 256]
        sv[f"ry0{h}"] = ry0 = pf_obj @ w0[:, 1024:].t() + cset.reshape(B, 2, 1, 256).expand(B, 2, N, 256).reshape(B, 2 * N, 256)
        sv[f"rst0{h}"] = st0 = _gn_stats(ry0.reshape(B, 2 * N, 32, 8), (1, 3))
        xh = ((ry0.reshape(B, 2 * N, 32, 8) - st0[0]) * st0[1]).reshape(B, 2 * N, 256)
        sv[f"ru0{h}"] = ru0 = F.gelu(xh * w[pre + ".layers.1.weight"] + w[pre + ".layers.1.bias"])
        sv[f"ry1{h}"] = ry1 = ru0 @ w[pre + ".layers.3.weight"][:, :, 0].t() + w[pre + ".layers.3.bias"]
        sv[f"rst1{h}"] = st1 = _gn_stats(ry1.reshape(B, 2 * N, 32, 8), (1, 3))
        xh = ((ry1.reshape(B, 2 * N, 32, 8) - st1[0]) * st1[1]).reshape(B, 2 * N, 256)
        ru1 = F.gelu(xh * w[pre + ".layers.4.weight"] + w[pre + ".layers.4.bias"])
        wp = w[pre + ".conv_p.weight"][0, :, 0]  # [P]
        sv[f"rwsum{h}"] = wsum = (ru1 * wp.reshape(1, -1, 1)).sum(dim=1)  # [B, 256]  (rot tail, by linearity)
        sv[f"rv{h}"] = None  # v_p = neck(u1_p) is only needed for dwp; recomputed in the backward
        r6.append(wsum @ w[pre + ".neck.0.weight"][:, :, 0].t() + w[pre + ".neck.0.bias"] * wp.sum() + w[pre + ".conv_p.bias"])
    sv["r6"] = r6 = torch.cat(r6, dim=1)
    sv["d_t"], sv["d_s"] = d_t, d_s
    rot, t, s = co.pose_update(co.rot6d_to_mat(r6), d_t, d_s, pose[:, :3, :3], pose[:, :3, 3], scale, K)
    sv["rot"], sv["t"], sv["s"] = rot, t, s
    return sv


def loss_backward(rot, trans, scale, gt_rot, gt_trans, gt_scale, kps, sym_info):
    """Gradients of the summed shipped losses w.r.t. the predicted (R, t, s): what the CUDA loss kernel emits."""
    B, n = kps.shape[0], kps.shape[1]
    gt_sym = closest_sym_rot(rot, gt_rot, sym_info)
    sk = kps * scale.unsqueeze(1)
    d_est = torch.sign(sk @ rot.transpose(1, 2) - (kps * gt_scale.unsqueeze(1)) @ gt_sym.transpose(1, 2)) / float(B * n)
    d_rot = d_est.transpose(1, 2) @ sk  # sum_p d_est_p (s*k_p)^T
    d_scale = ((d_est @ rot) * kps).sum(dim=1)
    is_sym = torch.tensor([s is not None for s in sym_info])
    n_sym, n_nosym = int(is_sym.sum()), int((~is_sym).sum())
    for b in range(B):
        if is_sym[b]:
            d_rot[b, :, 1] += torch.sign(rot[b, :, 1] - gt_rot[b, :, 1]) / (3.0 * n_sym)
        else:
            d_rot[b] += -gt_rot[b] / (4.0 * n_nosym)
    d_trans = torch.empty_like(trans)
    d_trans[:, :2] = torch.sign(trans[:, :2] - gt_trans[:, :2]) / (2.0 * B)
    d_trans[:, 2] = torch.sign(trans[:, 2] - gt_trans[:, 2]) / float(B)
    d_scale = d_scale + torch.sign(scale - gt_scale) / (3.0 * B)
    return d_rot, d_trans, d_scale


def pose_backward(sv, pose, K, d_rot, d_trans, d_scale):
    """Backward of the pose update (pose_scale_from_delta_init.py:47-95) and the rot6d Gram-Schmidt
    (core/utils/rot_reps.py:34-55) -> (d r6 [B,6], d Delta_t [B,3], d Delta_s [B,3])."""
    r_in, t_in = pose[:, :3, :3], pose[:, :3, 3]
    d_dr = d_rot @ r_in.transpose(1, 2)  # R' = dR . R
    fx, fy = K[:, 0, 0], K[:, 1, 1]
    dt = sv["d_t"]
    z_new = dt[:, 2] * t_in[:, 2]
    dz = d_trans[:, 2] + d_trans[:, 0] * (dt[:, 0] / fx + t_in[:, 0] / t_in[:, 2]) + d_trans[:, 1] * (dt[:, 1] / fy + t_in[:, 1] / t_in[:, 2])
    d_dt = torch.stack((d_trans[:, 0] * z_new / fx, d_trans[:, 1] * z_new / fy, dz * t_in[:, 2]), dim=1)
    # Gram-Schmidt: x = a/|a|, c = x x b, z = c/|c|, y = z x x, dR = [x y z] (columns)
    a, b = sv["r6"][:, :3], sv["r6"][:, 3:]
    na = a.norm(dim=1, keepdim=True)
    x = a / na
    c = torch.cross(x, b, dim=1)
    nc = c.norm(dim=1, keepdim=True)
    z = c / nc
    gx, gy, gz = d_dr[:, :, 0], d_dr[:, :, 1], d_dr[:, :, 2]
    gz = gz + torch.cross(x, gy, dim=1)  # y = z x x
    gx = gx + torch.cross(gy, z, dim=1)
    gc = (gz - z * (z * gz).sum(1, keepdim=True)) / nc
    gx = gx + torch.cross(b, gc, dim=1)  # c = x x b
    gb = torch.cross(gc, x, dim=1)
    ga = (gx - x * (x * gx).sum(1, keepdim=True)) / na
    return torch.cat((ga, gb), dim=1), d_dt, d_scale


def manual_backward(w: Weights, sv: dict, d_r6: torch.Tensor, d_dt: torch.Tensor, d_ds: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Parameter gradients of one iteration from the head-output gradients, stage by stage as the CUDA chain does."""
    q = sv["q"]
    S, N, _ = q.shape
    B, P = S // 2, 2 * N
    gr: Dict[str, torch.Tensor] = {}

    def lin_bwd(name, x, dy, conv):
        """dW = dy^T x (reduced over all rows), db = sum dy, returns dx = dy W."""
        wt = w[name + ".weight"][:, :, 0] if conv else w[name + ".weight"]
        dyf, xf = dy.reshape(-1, dy.shape[-1]), x.reshape(-1, x.shape[-1])
        gw = dyf.t() @ xf
        gr[name + ".weight"] = gr.get(name + ".weight", 0) + (gw.unsqueeze(2) if conv else gw)
        gr[name + ".bias"] = gr.get(name + ".bias", 0) + dyf.sum(0)
        return dy @ wt

    # ---- ts head (fc_trans_size_head.py:61-70), B rows
    du1 = lin_bwd("ts_head.fc_t", sv["ts_u1"], d_dt, False) + lin_bwd("ts_head.fc_s", sv["ts_u1"], d_ds, False)
    view8 = lambda t: t.reshape(B, 32, 8)
    dy1, dga, dbe = _gn_gelu_backward(view8(du1), view8(sv["ts_y1"]), *sv["ts_st1"], view8(w["ts_head.linears.4.weight"].expand(B, 256)),
                                      view8(w["ts_head.linears.4.bias"].expand(B, 256)), lambda t: t, (2,))
    gr["ts_head.linears.4.weight"], gr["ts_head.linears.4.bias"] = dga.reshape(B, 256).sum(0), dbe.reshape(B, 256).sum(0)
    du0 = lin_bwd("ts_head.linears.3", sv["ts_u0"], dy1.reshape(B, 256), False)
    dy0, dga, dbe = _gn_gelu_backward(view8(du0), view8(sv["ts_y0"]), *sv["ts_st0"], view8(w["ts_head.linears.1.weight"].expand(B, 256)),
                                      view8(w["ts_head.linears.1.bias"].expand(B, 256)), lambda t: t, (2,))
    gr["ts_head.linears.1.weight"], gr["ts_head.linears.1.bias"] = dga.reshape(B, 256).sum(0), dbe.reshape(B, 256).sum(0)
    d_in = lin_bwd("ts_head.linears.0", sv["ts_in"], dy0.reshape(B, 256), False)  # [B, 1091]; the init-scale part is detached
    dg = torch.zeros(S, 1024, dtype=q.dtype)
    dg[0::2] = d_in[:, :1024]
    dpfmax = torch.zeros(S, 64, dtype=q.dtype)
    dpfmax[0::2] = d_in[:, 1024:1088]

    # ---- rotation heads (conv_out_per_rot_head.py:126-140)
    dpf = torch.zeros(S, N, 64, dtype=q.dtype)
    pf_obj = sv["pf"].reshape(B, P, 64)
    for h, pre in enumerate(("rot_head.rot_head_x", "rot_head.rot_head_y")):
        dr = d_r6[:, 3 * h:3 * h + 3]  # [B, 3]
        wp = w[pre + ".conv_p.weight"][0, :, 0]
        wn, bn = w[pre + ".neck.0.weight"][:, :, 0], w[pre + ".neck.0.bias"]
        # recompute u1 (the GPU keeps y1 and the statistics)
        g1, b1 = w[pre + ".layers.4.weight"], w[pre + ".layers.4.bias"]
        v4 = lambda t: t.reshape(B, P, 32, 8)
        xh1 = ((v4(sv[f"ry1{h}"]) - sv[f"rst1{h}"][0]) * sv[f"rst1{h}"][1]).reshape(B, P, 256)
        u1 = F.gelu(xh1 * g1 + b1)
        # conv_p / neck: r_j = sum_p wp_p (Wn u1_p + bn)_j + bp
        e = dr @ wn  # [B, 256]: e_b = Wn^T dr_b, so du1[b, p, :] = wp_p e_b (rank one)
        gr[pre + ".conv_p.bias"] = dr.sum().reshape(1)
        gr[pre + ".conv_p.weight"] = ((u1 * e.unsqueeze(1)).sum(2) + (dr @ bn).unsqueeze(1)).sum(0).reshape(1, P, 1)
        gr[pre + ".neck.0.weight"] = (dr.t() @ sv[f"rwsum{h}"]).unsqueeze(2)
        gr[pre + ".neck.0.bias"] = dr.sum(0) * wp.sum()
        du1 = wp.reshape(1, P, 1) * e.unsqueeze(1)
        dy1, dga, dbe = _gn_gelu_backward(v4(du1), v4(sv[f"ry1{h}"]), *sv[f"rst1{h}"], v4(g1.expand(B, P, 256)), v4(b1.expand(B, P, 256)),
                                          lambda t: t, (1, 3))
        gr[pre + ".layers.4.weight"], gr[pre + ".layers.4.bias"] = dga.reshape(-1, 256).sum(0), dbe.reshape(-1, 256).sum(0)
        du0 = lin_bwd(pre + ".layers.3", sv[f"ru0{h}"], dy1.reshape(B, P, 256), True)
        g0, b0 = w[pre + ".layers.1.weight"], w[pre + ".layers.1.bias"]
        dy0, dga, dbe = _gn_gelu_backward(v4(du0), v4(sv[f"ry0{h}"]), *sv[f"rst0{h}"], v4(g0.expand(B, P, 256)), v4(b0.expand(B, P, 256)),
                                          lambda t: t, (1, 3))
        gr[pre + ".layers.1.weight"], gr[pre + ".layers.1.bias"] = dga.reshape(-1, 256).sum(0), dbe.reshape(-1, 256).sum(0)
        dy0 = dy0.reshape(B, P, 256)
        # layer 0, split: point-feature part per point, global-feature part once per set
        w0 = w[pre + ".layers.0.weight"][:, :, 0]
        dcset = dy0.reshape(S, N, 256).sum(1)  # [S, 256]
        gw0 = torch.cat((dcset.t() @ sv["g"], dy0.reshape(-1, 256).t() @ pf_obj.reshape(-1, 64)), dim=1)
        gr[pre + ".layers.0.weight"] = gw0.unsqueeze(2)
        gr[pre + ".layers.0.bias"] = dcset.sum(0)
        dg = dg + dcset @ w0[:, :1024]
        dpf = dpf + (dy0 @ w0[:, 1024:]).reshape(S, N, 64)

    # ---- encoder (pointnets/pointnet.py:97-116), all S sets at once (shared weights)
    def max_layer_bwd(name, x, arg, dmax, relu_max):
        """Sparse backward of `max over points of [relu](x W^T + b)`: only the arg-max point of each (set, channel)
        receives a gradient.  Returns dx [S, N, K] (non-zero on the arg-max rows only)."""
        wt = w[name + ".weight"][:, :, 0]  # [C, K]
        d = dmax if relu_max is None else dmax * (relu_max > 0)
        xs = torch.gather(x, 1, arg.unsqueeze(2).expand(-1, -1, x.shape[2]))  # [S, C, K]: the arg-max rows
        gr[name + ".weight"] = gr.get(name + ".weight", 0) + (d.unsqueeze(2) * xs).sum(0).unsqueeze(2)
        gr[name + ".bias"] = gr.get(name + ".bias", 0) + d.sum(0)
        dx = torch.zeros_like(x)
        dx.scatter_add_(1, arg.unsqueeze(2).expand(-1, -1, x.shape[2]), d.unsqueeze(2) * wt.unsqueeze(0))
        return dx

    def tnet_bwd(pre, x_in, tag, dt_mat):
        """Backward of a T-Net given d(T) [S, k, k]; returns d(x_in)."""
        d = lin_bwd(pre + ".fc3", sv[tag + "fc2"], dt_mat.reshape(S, -1), False) * (sv[tag + "fc2"] > 0)
        d = lin_bwd(pre + ".fc2", sv[tag + "fc1"], d, False) * (sv[tag + "fc1"] > 0)
        dmax = lin_bwd(pre + ".fc1", sv[tag + "max"], d, False)
        d128 = max_layer_bwd(pre + ".conv3", sv[tag + "128"], sv[tag + "arg"], dmax, sv[tag + "max"]) * (sv[tag + "128"] > 0)
        d64 = lin_bwd(pre + ".conv2", sv[tag + "64"], d128, True) * (sv[tag + "64"] > 0)
        return lin_bwd(pre + ".conv1", x_in, d64, True)

    dpf.scatter_add_(1, sv["pfarg"].unsqueeze(1), dpfmax.unsqueeze(1))  # max_n pointfeat of the ts-head input
    d512 = max_layer_bwd("pcl_net.conv4", sv["a512"], sv["garg"], dg, None) * (sv["a512"] > 0)
    d128 = lin_bwd("pcl_net.conv3", sv["a128"], d512, True) * (sv["a128"] > 0)
    dpf = dpf + lin_bwd("pcl_net.conv2", sv["pf"], d128, True)
    # pf = h1 . T64 per set
    dh1 = dpf @ sv["t64"].transpose(1, 2)
    dt64 = sv["h1"].transpose(1, 2) @ dpf  # [S, 64, 64]
    dh1 = dh1 + tnet_bwd("pcl_net.fstn", sv["h1"], "f", dt64)
    dqp = lin_bwd("pcl_net.conv1", sv["qp"], dh1 * (sv["h1"] > 0), True)
    dt3 = q.transpose(1, 2) @ dqp  # q' = q . T3 per set; the points themselves are inputs (no gradient)
    tnet_bwd("pcl_net.stn", q, "s", dt3)
    return gr


def manual_train_step(w, pcl, kps, pose, scale, K, gt_pose, gt_scale, sym_info):
    """train_step without autograd: manual_forward + loss_backward + pose_backward + manual_backward."""
    x, tfd = co.update_points(pcl, kps, pose, scale)
    B, N = pcl.shape[0], pcl.shape[1]
    q = torch.stack((x.permute(0, 2, 1), tfd.permute(0, 2, 1)), dim=1).reshape(2 * B, N, 3)
    sv = manual_forward(w, q, pose, scale, K)
    d_rot, d_trans, d_scale = loss_backward(sv["rot"], sv["t"], sv["s"], gt_pose[:, :3, :3], gt_pose[:, :3, 3], gt_scale, kps, sym_info)
    d_r6, d_dt, d_ds = pose_backward(sv, pose, K, d_rot, d_trans, d_scale)
    return sv, manual_backward(w, sv, d_r6, d_dt, d_ds)
