"""Per-stage parity: the engine's intermediate buffers after ONE refinement iteration, read through
the C ABI's catre_debug_read taps, against the oracle's intermediates of the same stage (reference lines in
oracle/catre_oracle.py).  The end-to-end tests only see (R, t, s); a stage that is wrong by a small factor can hide
behind the heads' normalisations, this test names it.

Tolerance per stage: REL * max(1, max|reference stage|); REL = 1e-5 (fp32 mode) / 3e-5 (f16x3: operands carry 22
significand bits, products are accumulated in fp32)."""
import os

import pytest
import torch
import torch.nn.functional as F

from catre_b200 import engine, synth
from oracle import catre_oracle as O

pytestmark = pytest.mark.gpu

REL = {"fp32": 1e-5, "f16x3": 3e-5}


def keys2f(k):
    k = k.clone()
    neg = k < 0
    k[neg] = k[neg] ^ 0x7FFFFFFF
    return k.view(torch.float32)


def pair16(eng, name, shape):
    """fp16 hi + fp16 residual operand pair -> fp32 value"""
    hi = eng.debug_read(name + "_hi", shape, torch.int16).view(torch.float16).float()
    lo = eng.debug_read(name + "_lo", shape, torch.int16).view(torch.float16).float()
    return hi + lo


@pytest.mark.parametrize("prec", ["fp32", "f16x3"])
def test_every_stage_matches_the_oracle(prec):
    B, N = 3, 1024
    S = 2 * B
    w = synth.load_weights()
    b = synth.make_batch(B, N, seed=5)
    os.environ["CATRE_DEBUG_TAPS"] = "1"  # read at catre_create: the tensor-core modes then keep a copy of the rot layer-1 output
    try:
        eng = engine.Engine(N, B, prec, 0)
    finally:
        del os.environ["CATRE_DEBUG_TAPS"]
    eng.load_weights(w)
    d = b.to("cuda")
    poses, scales = eng.refine(d.pcl, d.prior, d.init_pose, d.init_scale, d.K, 1)
    torch.cuda.synchronize()
    tc = prec != "fp32"
    report = []

    def check(name, got, ref):
        scale = max(1.0, ref.abs().max().item())
        err = (got.double().cpu() - ref.double()).abs().max().item()
        report.append(f"{name}:{err / scale:.1e}")
        assert err <= REL[prec] * scale, (prec, name, err, scale, report)

    # U1 (batch_test.py:85-97): sets ordered 2b = observed, 2b+1 = prior
    x, tfd = O.update_points(b.pcl, b.prior, b.init_pose, b.init_scale)
    q = torch.stack((x, tfd), dim=1).reshape(S, 3, N)
    check("q", eng.debug_read("q", (S, N, 3)), q.permute(0, 2, 1))
    # E1 STN3d (pointnet.py:24-41)
    h = F.relu(O._pw(w, "pcl_net.stn.conv1", q))
    h = F.relu(O._pw(w, "pcl_net.stn.conv2", h))
    h = F.relu(O._pw(w, "pcl_net.stn.conv3", h))
    check("gmax_stn", keys2f(eng.debug_read("gmax_stn", (S, 1024), torch.int32)), h.max(2)[0])
    t3 = O.tnet(w, "pcl_net.stn", q, 3)
    check("t3", eng.debug_read("t3", (S, 9)), t3.reshape(S, 9))
    # E2 input transform + conv1 (pointnet.py:100-103)
    xq = torch.bmm(q.transpose(2, 1), t3).transpose(2, 1)
    h1 = F.relu(O._pw(w, "pcl_net.conv1", xq))
    h1_got = pair16(eng, "x64", (S, N, 64)) if tc else eng.debug_read("h64a", (S, N, 64))
    check("h1", h1_got, h1.permute(0, 2, 1))
    # E3 STNkd (pointnet.py:57-78); the engine keeps T64^T
    hf = F.relu(O._pw(w, "pcl_net.fstn.conv1", h1))
    if tc:
        check("fstn_conv1", pair16(eng, "f64", (S, N, 64)), hf.permute(0, 2, 1))
    hf = F.relu(O._pw(w, "pcl_net.fstn.conv2", hf))
    hf = F.relu(O._pw(w, "pcl_net.fstn.conv3", hf))
    check("gmax_fstn", keys2f(eng.debug_read("gmax_fstn", (S, 1024), torch.int32)), hf.max(2)[0])
    t64 = O.tnet(w, "pcl_net.fstn", h1, 64)
    t64_got = pair16(eng, "t64s", (S, 4096)) if tc else eng.debug_read("t64", (S, 4096))
    check("t64", t64_got, t64.transpose(1, 2).reshape(S, 4096))
    # E4 feature transform, trunk, global max (pointnet.py:105-116)
    pf = torch.bmm(h1.transpose(2, 1), t64).transpose(2, 1)
    pf_got = pair16(eng, "pf", (S, N, 64)) if tc else eng.debug_read("h64b", (S, N, 64))
    check("pf", pf_got, pf.permute(0, 2, 1))
    check("gmax_pf", keys2f(eng.debug_read("gmax_pf", (S, 64), torch.int32)), pf.max(2)[0])
    a2 = F.relu(O._pw(w, "pcl_net.conv2", pf))
    if tc:
        check("conv2", pair16(eng, "a128", (S, N, 128)), a2.permute(0, 2, 1))
    g, _ = O.pointnet_feat(w, q)
    check("gmax_g", keys2f(eng.debug_read("gmax_g", (S, 1024), torch.int32)), g)
    # R1 layer-0 split: cset = W0[:, :1024] . g_set + b0 (both heads stacked), and ts layer 0 over the observed g
    w0 = torch.cat([w[f"rot_head.rot_head_{a}.layers.0.weight"][:, :1024, 0] for a in "xy"], 0)
    b0 = torch.cat([w[f"rot_head.rot_head_{a}.layers.0.bias"] for a in "xy"], 0)
    check("cset", eng.debug_read("cset", (S, 512)), g @ w0.T + b0)
    wt0 = w["ts_head.linears.0.weight"]
    check("ts0", eng.debug_read("ts0", (B, 256)), g[0::2] @ wt0[:, :1024].T + w["ts_head.linears.0.bias"])
    # R1 layer 1 output (+ bias), both heads: fp32 in every mode
    feat = torch.cat((g.unsqueeze(2).expand(-1, -1, N), pf), dim=1)
    rot_feat = feat.reshape(B, 2, 1088, N).permute(0, 2, 1, 3).reshape(B, 1088, 2 * N)
    u = []
    for a in "xy":
        pre = f"rot_head.rot_head_{a}"
        a0 = O._pw(w, pre + ".layers.0", rot_feat)
        u0 = F.gelu(F.group_norm(a0, 32, w[pre + ".layers.1.weight"], w[pre + ".layers.1.bias"], 1e-5))
        u.append(O._pw(w, pre + ".layers.3", u0))
    a1 = torch.cat(u, dim=1)  # [B, 512, P]
    if tc:  # a1T [B][P/4][512][4]
        got = eng.debug_read("a1", (B, 2 * N // 4, 512, 4)).permute(0, 2, 1, 3).reshape(B, 512, 2 * N)
    else:
        got = eng.debug_read("a1", (B, 2 * N, 512)).permute(0, 2, 1)
    check("rot_layer1", got, a1)
    rp, rs = O.refine(w, b.pcl, b.prior, b.init_pose, b.init_scale, b.K, 1)
    check("pose", poses.cpu(), rp)
    check("scale", scales.cpu(), rs)
    print(f"\n[stage parity {prec}] relative errors: " + " ".join(report))
    eng.close()
