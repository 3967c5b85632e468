"""Stage-by-stage comparison of the engine's internal buffers against the oracle (debug aid)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catre_b200 import engine, synth  # noqa: E402
from oracle import catre_oracle as O  # noqa: E402


def keys2f(k):
    k = k.clone()
    neg = k < 0
    k[neg] = k[neg] ^ 0x7FFFFFFF
    return k.view(torch.float32)


def main():
    prec = "fp32"  # the tensor-core modes keep only bf16 hi/lo operands; the fp32 taps below exist in fp32 mode
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    N = 1024
    w = synth.load_weights()
    b = synth.make_batch(B, N, seed=5)
    eng = engine.Engine(N, B, prec, 0)
    eng.load_weights(w)
    d = b.to("cuda")
    poses, scales = eng.refine(d.pcl, d.prior, d.init_pose, d.init_scale, d.K, 1)
    torch.cuda.synchronize()
    S = 2 * B
    # oracle intermediates, sets ordered (2b = obs, 2b+1 = prior)
    x, tfd = O.update_points(b.pcl, b.prior, b.init_pose, b.init_scale)
    q = torch.stack((x, tfd), dim=1).reshape(S, 3, N)

    def rep(name, got, ref):
        err = (got.double() - ref.double()).abs().max().item()
        print(f"{name:12s} max|err| {err:.3e}   ref max {ref.abs().max().item():.3e}", flush=True)

    rep("q", eng.debug_read("q", (S, N, 3)), q.permute(0, 2, 1))
    h = F.relu(O._pw(w, "pcl_net.stn.conv1", q))
    h = F.relu(O._pw(w, "pcl_net.stn.conv2", h))
    h = F.relu(O._pw(w, "pcl_net.stn.conv3", h))
    rep("gmax_stn", keys2f(eng.debug_read("gmax_stn", (S, 1024), torch.int32)), h.max(2)[0])
    t3 = O.tnet(w, "pcl_net.stn", q, 3)
    rep("t3", eng.debug_read("t3", (S, 9)), t3.reshape(S, 9))
    xq = torch.bmm(q.transpose(2, 1), t3).transpose(2, 1)
    h1 = F.relu(O._pw(w, "pcl_net.conv1", xq))
    rep("h1", eng.debug_read("h64a", (S, N, 64)), h1.permute(0, 2, 1))
    t64 = O.tnet(w, "pcl_net.fstn", h1, 64)
    rep("t64", eng.debug_read("t64", (S, 4096)), t64.transpose(1, 2).reshape(S, 4096))  # the engine keeps T64^T
    pf = torch.bmm(h1.transpose(2, 1), t64).transpose(2, 1)
    rep("pf", eng.debug_read("h64b", (S, N, 64)), pf.permute(0, 2, 1))
    rep("gmax_pf", keys2f(eng.debug_read("gmax_pf", (S, 64), torch.int32)), pf.max(2)[0])
    g, _ = O.pointnet_feat(w, q)
    rep("gmax_g", keys2f(eng.debug_read("gmax_g", (S, 1024), torch.int32)), g)
    # rot head layer 0 (both heads stacked)
    feat = torch.cat((g.unsqueeze(2).expand(-1, -1, N), pf), dim=1)  # [S,1088,N]
    rot_feat = feat.reshape(B, 2, 1088, N).permute(0, 2, 1, 3).reshape(B, 1088, 2 * N)
    a0 = torch.cat([O._pw(w, f"rot_head.rot_head_{a}.layers.0", rot_feat) for a in "xy"], dim=1)  # [B,512,P]
    rep("a0", eng.debug_read("a0", (B, 2 * N, 512)), a0.permute(0, 2, 1))
    u = []
    for i, a in enumerate("xy"):
        pre = f"rot_head.rot_head_{a}"
        u0 = F.gelu(F.group_norm(a0[:, i * 256:(i + 1) * 256], 32, w[pre + ".layers.1.weight"], w[pre + ".layers.1.bias"], 1e-5))
        u.append(O._pw(w, pre + ".layers.3", u0))
    a1 = torch.cat(u, dim=1)
    rep("a1", eng.debug_read("a1", (B, 2 * N, 512)), a1.permute(0, 2, 1))
    rp, rs = O.refine(w, b.pcl, b.prior, b.init_pose, b.init_scale, b.K, 1)
    rep("pose", poses.cpu(), rp)
    rep("scale", scales.cpu(), rs)


if __name__ == "__main__":
    main()
