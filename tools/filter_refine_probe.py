import os, sys, torch, torch.nn.functional as F
sys.path.insert(0, "/root/repo")
from catre_b200 import synth
from oracle import catre_oracle as orc
torch.set_num_threads(os.cpu_count())
w = synth.load_weights()
bt = synth.make_batch(8, 1024, 5)
cap = {}
orig = orc._pw
def _pw(wt, name, x):
    if name in ("pcl_net.conv4", "pcl_net.stn.conv3", "pcl_net.fstn.conv3"):
        cap.setdefault(name, []).append(x.clone())
    return orig(wt, name, x)
orc._pw = _pw
orc.refine(w, bt.pcl, bt.prior, bt.init_pose, bt.init_scale, bt.K, 2)
for name, xs in cap.items():
    x = torch.cat(xs[-2:], 0)  # last iteration: obs sets + prior sets  [S, K, N]
    W = w[name + ".weight"][:, :, 0]; b = w[name + ".bias"]
    S, K, N = x.shape
    f = torch.einsum("ck,skn->scn", W.double(), x.double())
    xh = x.half().float(); Wh = W.half().float()
    ft = torch.einsum("ck,skn->scn", Wh, xh).double()
    err = (f - ft).abs()
    wn2 = W.double().norm(dim=1); an2 = x.double().norm(dim=1)  # [C], [S,N]
    bA = 2.0**-11 * wn2[None, :, None] * an2[:, None, :]
    bB = 2.0**-11 * torch.einsum("ck,skn->scn", W.double().abs(), x.double().abs())
    print(name, "K", K, "f range", f.min().item(), f.max().item(), "err max", err.max().item(), "err rms", err.pow(2).mean().sqrt().item(),
          "bound A mean", bA.mean().item(), "bound B mean", bB.mean().item(), "max err/bA", (err / bA).max().item(), "max err/bB", (err/bB).max().item())
    # candidates per (set, channel, 256-pt tile) ; and per set (global max known)
    for T in (256, 1024):
        ftt = ft.reshape(S, -1, N // T, T)
        for nm, bd in (("A", bA), ("B", bB), ("A/4", bA / 4), ("A/8", bA/8)):
            bdt = bd.reshape(S, -1, N // T, T)
            # per tile eps = max over pts in tile of bound (per channel)
            eps = bdt.max(dim=3, keepdim=True)[0]
            m = ftt.max(dim=3, keepdim=True)[0]
            cand = (ftt >= m - 2 * eps).sum(dim=3).double()
            # pointwise eps variant: candidate if ft + eps_p >= max_p (ft - eps_p)
            lo = (ftt - bdt).max(dim=3, keepdim=True)[0]
            cand2 = ((ftt + bdt) >= lo).sum(dim=3).double()
            print(f"   tile {T} bound {nm}: cand/(ch,tile) tile-eps mean {cand.mean().item():.2f} max {cand.max().item():.0f} | point-eps mean {cand2.mean().item():.2f} max {cand2.max().item():.0f} p99 {cand2.flatten().quantile(0.99).item():.0f}")
    # distinct candidate points per set, global criterion with bound A (point-eps)
    lo = (ft - bA).max(dim=2, keepdim=True)[0]
    candm = (ft + bA) >= lo        # [S, C, N]
    pts = candm.any(dim=1).sum(dim=1).double()
    print(f"   global criterion bound A: candidates/(set,ch) mean {candm.sum(2).double().mean():.2f}; distinct candidate points per set mean {pts.mean():.0f} max {pts.max():.0f} of {N}")
    # running criterion: tiles of 256 in order, threshold = max(tile max, running max so far)
    T = 256
    run = torch.full((S, f.shape[1], 1), -1e30, dtype=torch.float64)
    tot = 0; ptsets = torch.zeros(S, N, dtype=torch.bool)
    for t0 in range(0, N, T):
        ftt = ft[:, :, t0:t0+T]; bt_ = bA[:, :, t0:t0+T]
        lo_t = torch.maximum((ftt - bt_).max(dim=2, keepdim=True)[0], run)
        cm = (ftt + bt_) >= lo_t
        tot += cm.sum().item(); ptsets[:, t0:t0+T] |= cm.any(dim=1)
        run = lo_t
    print(f"   running criterion (256-pt tiles in order): candidates/(set,ch) {tot / (S * f.shape[1]):.2f}; distinct points per set mean {ptsets.sum(1).double().mean():.0f} max {ptsets.sum(1).max()}")
